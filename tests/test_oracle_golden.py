"""The CPU oracle against the golden vectors produced by the reference itself (tests/golden/make_golden.py)."""

import pytest

import oracle
from tests import _golden as G


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_build(name):
    G.check_face_tree_build(oracle.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_points_and_weights(name):
    G.check_face_tree_points(oracle.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_boxes(name):
    G.check_face_tree_boxes(oracle.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_edges(name):
    G.check_face_tree_edges(oracle.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_faces(name):
    G.check_face_tree_faces(oracle.CellTree2d, name)


@pytest.mark.parametrize("name", G.EDGE_CASES)
def test_edge_tree(name):
    G.check_edge_tree(oracle.EdgeCellTree2d, name)


@pytest.mark.filterwarnings("ignore:overflow encountered", "ignore:invalid value encountered")  # t of a hit on an overflowing segment
def test_extreme_segments():
    G.check_extreme_segments(oracle.CellTree2d, oracle.EdgeCellTree2d)
