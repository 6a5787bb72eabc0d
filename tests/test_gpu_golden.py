"""
The CUDA path (through the C-ABI, via the drop-in classes) against the golden vectors produced by the
reference itself.  Same checkers as tests/test_oracle_golden.py: indices, pair order, nodes and bb_indices
bit-exact; floats within 1e-12 relative (and in fact bit-equal).
"""

import pytest

from tests import _golden as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import numba_celltree_b200

    return numba_celltree_b200


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_build(pkg, name):
    G.check_face_tree_build(pkg.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_points_and_weights(pkg, name):
    G.check_face_tree_points(pkg.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_boxes(pkg, name):
    G.check_face_tree_boxes(pkg.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_edges(pkg, name):
    G.check_face_tree_edges(pkg.CellTree2d, name)


@pytest.mark.parametrize("name", G.FACE_CASES)
def test_faces(pkg, name):
    G.check_face_tree_faces(pkg.CellTree2d, name)


@pytest.mark.parametrize("name", G.EDGE_CASES)
def test_edge_tree(pkg, name):
    G.check_edge_tree(pkg.EdgeCellTree2d, name)


@pytest.mark.filterwarnings("ignore:overflow encountered", "ignore:invalid value encountered")  # t of a hit on an overflowing segment
def test_extreme_segments(pkg):
    G.check_extreme_segments(pkg.CellTree2d, pkg.EdgeCellTree2d)
