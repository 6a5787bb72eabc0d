"""The CUDA path against the reference's own known-answer tests (restated in tests/_reference_kats.py)."""

import numpy as np
import pytest

from tests import _golden, _reference_kats as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import numba_celltree_b200

    return numba_celltree_b200


def test_init_and_casting(pkg):
    K.check_init_and_casting(pkg.CellTree2d)


def test_errors(pkg):
    K.check_errors(pkg.CellTree2d)


def test_point_lookups(pkg):
    K.check_point_lookups(pkg.CellTree2d)


def test_box_and_edge_lookup(pkg):
    K.check_box_and_edge_lookup(pkg.CellTree2d)


def test_example_material(pkg):
    K.check_example_material(pkg.CellTree2d)


def test_barycentric(pkg):
    K.check_barycentric(pkg.CellTree2d)


def test_locate_point_on_edge(pkg):
    K.check_locate_point_on_edge(pkg.CellTree2d)


def test_diagnostics(pkg):
    K.check_diagnostics(pkg.CellTree2d, _golden.load)


def test_edge_tree(pkg):
    from numba_celltree_b200.constants import CellTreeData

    K.check_edge_tree(pkg.EdgeCellTree2d, CellTreeData)


def test_locate_faces_rewrites_query_faces_in_place(pkg):
    # celltree.py:212: counter_clockwise is applied to the caller's query faces
    g = _golden.load("lattice_faces")
    tree = pkg.CellTree2d(g["vertices"], g["faces"], -1)
    faces = np.array([[3, 2, 1, 0, -1]], dtype=np.intp)  # clockwise unit quad on vertices 0..3
    vertices = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    tree.locate_faces(vertices, faces)
    assert faces.tolist() == [[0, 1, 2, 3, -1]]
