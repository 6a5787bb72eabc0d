"""
Golden fixture `extreme_segments.npz`: query segments with infinite, NaN, overflowing, subnormal, signed-zero and
axis-parallel coordinates, run through the REFERENCE's `intersect_edges` of both tree kinds (Deltares/numba_celltree
v0.4.2 at /root/reference, under Numba, in the build container; same conventions as make_golden.py).

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden_extreme.py

These inputs pin what the plane test of `locate_edge` (query.py:407-440) does outside the finite case, which is where
the CUDA path's division-free form of that test falls back to the reference's quotients (csrc/traverse.cuh).
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
from numba_celltree_b200.synthetic import delaunay_mesh, random_network  # noqa: E402

SPECIAL = np.array([np.inf, -np.inf, np.nan, 1e308, -1e308, 1.7e308, 5e-324, -5e-324, 0.0, -0.0, 1e-300, 0.5, 2.0])


def extreme_segments(rng, vertices, n=1500):
    a = rng.uniform(-0.2, 1.2, (n, 2))
    b = rng.uniform(-0.2, 1.2, (n, 2))
    which = rng.integers(0, 4, n)
    value = SPECIAL[rng.integers(0, len(SPECIAL), n)]
    a[which == 0, 0] = value[which == 0]
    a[which == 1, 1] = value[which == 1]
    b[which == 2, 0] = value[which == 2]
    b[which == 3, 1] = value[which == 3]
    both = rng.random(n) < 0.3
    b[both, 0] = SPECIAL[rng.integers(0, len(SPECIAL), int(both.sum()))]
    edges = np.stack((a, b), axis=1)
    axis = rng.uniform(0, 1, (600, 2, 2))
    axis[:300, 1, 0] = axis[:300, 0, 0]  # vertical
    axis[300:, 1, 1] = axis[300:, 0, 1]  # horizontal
    on_vertices = np.stack((vertices[rng.integers(0, len(vertices), 300)], vertices[rng.integers(0, len(vertices), 300)]), axis=1)
    degenerate = np.repeat(on_vertices[:20, :1], 2, axis=1)  # zero-length, on a vertex
    return np.concatenate([edges, axis, on_vertices, on_vertices[:, ::-1], degenerate])


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    vertices, faces = delaunay_mesh(2500, seed=99)
    segments = extreme_segments(rng, vertices)
    tree = CellTree2d(vertices, faces, -1)
    i, j, xy = tree.intersect_edges(segments)
    out.update(face_vertices=vertices, face_faces=faces, face_segments=segments, face_i=i, face_j=j, face_xy=xy)
    net_vertices, net_edges = random_network(1200, seed=4)
    lo, hi = net_vertices.min(), net_vertices.max()
    scaled = extreme_segments(rng, (net_vertices - lo) / (hi - lo))
    ordinary = np.isfinite(scaled) & (np.abs(scaled) < 1e3)
    scaled[ordinary] = lo + scaled[ordinary] * (hi - lo)
    net = EdgeCellTree2d(net_vertices, net_edges)
    i, j, xy = net.intersect_edges(scaled)
    out.update(net_vertices=net_vertices, net_edges=net_edges, net_segments=scaled, net_i=i, net_j=j, net_xy=xy)
    np.savez_compressed(HERE / "extreme_segments.npz", **out)
    print("extreme_segments.npz:", len(out["face_i"]), "face pairs,", len(out["net_i"]), "network pairs")


if __name__ == "__main__":
    main()
