"""
Golden tree arrays for two corners of creation.initialize that the other fixtures do not reach, from the REFERENCE:

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden_build_corners.py

  signed zeros   a quad mesh through the origin whose vertices on the axes carry -0.0 and +0.0 in alternation: the
                 reference's strict comparisons keep whichever zero comes first in a node's slice (get_bounds,
                 creation.py:153-171), and that zero -- sign included -- is what nodes["Lmax"] / nodes["Rmin"] hold.
                 The node arrays are stored as raw bytes so that the sign bit is compared.
  many buckets   n_buckets = 300 and 1000 (the reference accepts any n_buckets >= 2, celltree.py:69-72).
"""

import os
import pathlib
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, os.environ.get("CELLTREE_REFERENCE", "/root/reference"))

import numpy as np  # noqa: E402
from numba_celltree import CellTree2d, EdgeCellTree2d  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh  # noqa: E402


def signed_zero_mesh(nx, ny, seed):
    v, f = quad_mesh(nx, ny)
    v = v * [nx, ny] - [nx // 2, ny // 2]  # integer lattice through the origin
    rng = np.random.default_rng(seed)
    flip = rng.random(len(v)) < 0.5
    v[(v[:, 0] == 0) & flip, 0] = -0.0
    flip = rng.random(len(v)) < 0.5
    v[(v[:, 1] == 0) & flip, 1] = -0.0
    return v, f


def main():
    out = {}
    v, f = signed_zero_mesh(12, 10, 5)
    out["sz_vertices_bytes"] = np.frombuffer(v.tobytes(), dtype=np.uint8)
    out["sz_faces"] = f
    assert np.signbit(v[v == 0]).any() and not np.signbit(v[v == 0]).all()
    for nb_, cpl in [(4, 2), (2, 1), (3, 1), (8, 3), (2, 2)]:
        t = CellTree2d(v, f, -1, n_buckets=nb_, cells_per_leaf=cpl)
        out[f"sz_b{nb_}_c{cpl}_nodes_bytes"] = np.frombuffer(t.nodes.tobytes(), dtype=np.uint8)
        out[f"sz_b{nb_}_c{cpl}_bb_indices"] = t.bb_indices
        zeros = np.concatenate([t.nodes["Lmax"][t.nodes["Lmax"] == 0], t.nodes["Rmin"][t.nodes["Rmin"] == 0]])
        print("signed zero", nb_, cpl, "zero planes:", len(zeros), "negative:", int(np.signbit(zeros).sum()))
    # an edge network along the axes: zero-width boxes padded by the tolerance, plenty of exact zeros
    ev = np.array([[x, 0.0] for x in range(-6, 7)] + [[0.0, y] for y in range(-6, 7) if y != 0])
    ev[::2][ev[::2] == 0] = -0.0
    ee = np.array([[i, i + 1] for i in range(12)] + [[13 + i, 14 + i] for i in range(11)], dtype=np.int64)
    et = EdgeCellTree2d(ev, ee, n_buckets=4, cells_per_leaf=2)
    out["sz_edge_vertices_bytes"] = np.frombuffer(ev.tobytes(), dtype=np.uint8)
    out["sz_edge_edges"] = ee
    out["sz_edge_nodes_bytes"] = np.frombuffer(et.nodes.tobytes(), dtype=np.uint8)
    out["sz_edge_bb_indices"] = et.bb_indices

    dv, df = delaunay_mesh(4000, seed=21)
    out["mb_vertices"], out["mb_faces"] = dv, df
    for nb_, cpl in [(300, 2), (1000, 1), (70000 // 1000 * 16, 5)]:
        t = CellTree2d(dv, df, -1, n_buckets=nb_, cells_per_leaf=cpl)
        out[f"mb_b{nb_}_c{cpl}_nodes"] = t.nodes
        out[f"mb_b{nb_}_c{cpl}_bb_indices"] = t.bb_indices
        print("many buckets", nb_, cpl, "nodes:", len(t.nodes))
    pts = np.random.default_rng(3).uniform(0, 1, (5000, 2))
    out["mb_points"] = pts
    out["mb_locate_points_b300"] = CellTree2d(dv, df, -1, n_buckets=300).locate_points(pts)
    np.savez_compressed(HERE / "build_corners.npz", **out)
    print("wrote", HERE / "build_corners.npz")


if __name__ == "__main__":
    main()
