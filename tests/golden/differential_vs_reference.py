"""
Differential run of the CPU oracle against the REFERENCE itself (Deltares/numba_celltree v0.4.2 at /root/reference, under
Numba) on seeded inputs far larger than the committed fixtures.  Runs in the build container only (the reference does
not travel to the GPU box); prints one line per check and exits non-zero on the first difference.

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/differential_vs_reference.py [n_mesh_points] [seeds...]
"""

import os
import pathlib
import sys
import warnings

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
sys.path.insert(0, os.environ.get("CELLTREE_REFERENCE", "/root/reference"))
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import numba_celltree as nct  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
import oracle  # noqa: E402
from numba_celltree_b200.synthetic import c3_boxes, c4_edges, delaunay_mesh, quad_mesh, random_network  # noqa: E402


def same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and bool(((a == b) | ((a != a) & (b != b))).all())


def check(name, got, want):
    ok = all(same(np.asarray(g), np.asarray(w)) for g, w in zip(got, want)) and len(got) == len(want)
    print(f"{'ok  ' if ok else 'DIFF'} {name}: {', '.join(str(np.asarray(w).shape) for w in want)}", flush=True)
    if not ok:
        sys.exit(1)


def nodes_tuple(t):
    return tuple(t.nodes[f] for f in ("child", "Lmax", "Rmin", "ptr", "size", "dim")) + (t.bb_indices, t.bb_coords)


def main():
    n_mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    seeds = [int(s) for s in sys.argv[2:]] or [1, 2]
    for seed in seeds:
        rng = np.random.default_rng(seed)
        vertices, faces = delaunay_mesh(n_mesh, seed=seed)
        for n_buckets, cells_per_leaf in ((4, 2), (2, 1), (7, 3)):
            r = nct.CellTree2d(vertices, faces, -1, n_buckets=n_buckets, cells_per_leaf=cells_per_leaf)
            o = oracle.CellTree2d(vertices, faces, -1, n_buckets=n_buckets, cells_per_leaf=cells_per_leaf)
            check(f"seed {seed} delaunay({n_mesh}) build nb={n_buckets} cpl={cells_per_leaf}", nodes_tuple(o), nodes_tuple(r))
        points = rng.uniform(-0.02, 1.02, (1_000_000, 2))
        for tol in (None, 1e-7):
            check(f"seed {seed} locate_points + weights tol={tol}", o.compute_barycentric_weights(points, tol), r.compute_barycentric_weights(points, tol))
        boxes = c3_boxes(len(faces), 200_000)
        check(f"seed {seed} locate_boxes", o.locate_boxes(boxes), r.locate_boxes(boxes))
        check(f"seed {seed} intersect_boxes", o.intersect_boxes(boxes), r.intersect_boxes(boxes))
        edges = c4_edges(len(faces), 100_000)
        check(f"seed {seed} intersect_edges", o.intersect_edges(edges), r.intersect_edges(edges))
        side = max(8, int(np.sqrt(n_mesh / 2)))
        qv, qf = quad_mesh(side, side - 3)
        check(f"seed {seed} locate_faces", o.locate_faces(qv, qf.copy()), r.locate_faces(qv, qf.copy()))
        check(f"seed {seed} intersect_faces", o.intersect_faces(qv, qf, -1), r.intersect_faces(qv, qf, -1))
        # structured quads, queries on and next to the grid lines
        gv, gf = quad_mesh(300, 200)
        r, o = nct.CellTree2d(gv, gf, -1), oracle.CellTree2d(gv, gf, -1)
        check(f"seed {seed} quads(300x200) build", nodes_tuple(o), nodes_tuple(r))
        on_lines = gv[rng.integers(0, len(gv), 200_000)] + rng.choice([0.0, 1e-16, -1e-16, 1e-13, 0.0017], (200_000, 2))
        for tol in (None, 1e-9):
            check(f"seed {seed} quads points on lines tol={tol}", o.compute_barycentric_weights(on_lines, tol), r.compute_barycentric_weights(on_lines, tol))
        # tolerance 0: a point exactly on an edge makes the reference's Wachspress weights divide by zero, which Numba's
        # Python error model turns into ZeroDivisionError (barycentric_wachspress.py:76); only the lookup is compared
        check(f"seed {seed} quads points on lines tol=0.0 (lookup)", (o.locate_points(on_lines, 0.0),), (r.locate_points(on_lines, 0.0),))
        # network
        nv, ne = random_network(max(200, n_mesh // 10), seed=seed)
        r, o = nct.EdgeCellTree2d(nv, ne), oracle.EdgeCellTree2d(nv, ne)
        check(f"seed {seed} network build", nodes_tuple(o), nodes_tuple(r))
        a = rng.uniform(nv.min(0), nv.max(0), (100_000, 2))
        segs = np.stack((a, a + rng.normal(0, 3.0, a.shape)), axis=1)
        check(f"seed {seed} network intersect_edges", o.intersect_edges(segs), r.intersect_edges(segs))
        mids = 0.5 * (nv[ne[:, 0]] + nv[ne[:, 1]])
        check(f"seed {seed} network locate_points", (o.locate_points(np.concatenate([mids, a])),), (r.locate_points(np.concatenate([mids, a])),))
    print("oracle == reference on every check")


if __name__ == "__main__":
    main()
