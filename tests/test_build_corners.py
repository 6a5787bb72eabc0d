"""Tree construction in two corners the reference allows: bounding boxes that hold -0.0 beside +0.0 (the zero met first
in a node's slice is the one stored, sign included) and many buckets (n_buckets well above 255).  Golden arrays from the
reference (tests/golden/make_golden_build_corners.py); node arrays are compared as raw bytes."""

import pathlib

import numpy as np
import pytest

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "build_corners.npz"
SZ_PARAMS = [(4, 2), (2, 1), (3, 1), (8, 3), (2, 2)]
MB_PARAMS = [(300, 2), (1000, 1), (1120, 5)]


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def node_bytes(tree):
    return np.frombuffer(np.ascontiguousarray(tree.nodes).tobytes(), dtype=np.uint8)


def check(module, g):
    v = np.frombuffer(g["sz_vertices_bytes"].tobytes(), dtype=np.float64).reshape(-1, 2)
    for nb, cpl in SZ_PARAMS:
        t = module.CellTree2d(v, g["sz_faces"], -1, n_buckets=nb, cells_per_leaf=cpl)
        assert np.array_equal(t.bb_indices, g[f"sz_b{nb}_c{cpl}_bb_indices"])
        assert np.array_equal(node_bytes(t), g[f"sz_b{nb}_c{cpl}_nodes_bytes"]), (nb, cpl)
    ev = np.frombuffer(g["sz_edge_vertices_bytes"].tobytes(), dtype=np.float64).reshape(-1, 2)
    et = module.EdgeCellTree2d(ev, g["sz_edge_edges"], n_buckets=4, cells_per_leaf=2)
    assert np.array_equal(et.bb_indices, g["sz_edge_bb_indices"])
    assert np.array_equal(node_bytes(et), g["sz_edge_nodes_bytes"])
    for nb, cpl in MB_PARAMS:
        t = module.CellTree2d(g["mb_vertices"], g["mb_faces"], -1, n_buckets=nb, cells_per_leaf=cpl)
        assert np.array_equal(t.bb_indices, g[f"mb_b{nb}_c{cpl}_bb_indices"])
        assert np.array_equal(node_bytes(t), np.frombuffer(g[f"mb_b{nb}_c{cpl}_nodes"].tobytes(), dtype=np.uint8)), (nb, cpl)
    t = module.CellTree2d(g["mb_vertices"], g["mb_faces"], -1, n_buckets=300)
    assert np.array_equal(t.locate_points(g["mb_points"]), g["mb_locate_points_b300"])


def test_oracle_reproduces_the_reference(g):
    import oracle

    check(oracle, g)


@pytest.mark.gpu
def test_library_reproduces_the_reference(g):
    import numba_celltree_b200

    check(numba_celltree_b200, g)


@pytest.mark.gpu
def test_bucket_count_limits():
    from numba_celltree_b200 import CellTree2d
    from numba_celltree_b200.synthetic import quad_mesh

    v, f = quad_mesh(6, 5)
    CellTree2d(v, f, -1, n_buckets=65535)
    with pytest.raises(ValueError):
        CellTree2d(v, f, -1, n_buckets=65536)
