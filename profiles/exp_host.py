"""Experiment: calls with plain (pageable) NumPy inputs: constructor, locate_points, intersect_faces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh
def best(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r
v, f = quad_mesh(4096, 4096)
t, tree = best(lambda: CellTree2d(v, f, -1), reps=2)
print(f"CellTree2d(16.7 M quads) from NumPy arrays: {t*1e3:.1f} ms wall (device build {tree.build_ms:.1f} ms)")
pts = np.random.default_rng(42).uniform(0, 1, (50_000_000, 2))
t, r = best(lambda: tree.locate_points(pts))
print(f"locate_points, 50 M pageable points -> recycled pinned result: {t*1e3:.1f} ms = {50e6/t/1e9:.2f} Gq/s")
del tree, pts, r
v, f = delaunay_mesh(1_000_000, 1234)
tree = CellTree2d(v, f, -1)
qv, qf = quad_mesh(1000, 1000)
t, r = best(lambda: tree.intersect_faces(qv, qf, -1), reps=5)
print(f"intersect_faces C5, NumPy in / out: {t*1e3:.2f} ms = {len(r[0])/t/1e6:.0f} Mpairs/s")
