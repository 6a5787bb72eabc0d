"""Experiment: intersect_faces (C5), NumPy in / NumPy out, with and without page-locked result arrays."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh
v, f = delaunay_mesh(1_000_000, 1234)
tree = CellTree2d(v, f, -1)
qv, qf = quad_mesh(1000, 1000)
for pinned in (False, True, True):
    _lib.set_pinned_results(pinned)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        i, j, a = tree.intersect_faces(qv, qf, -1)
        ts.append(time.perf_counter() - t0)
        n, area = len(i), float(a.sum())
        del i, j, a
    print(f"pinned={pinned}: " + " ".join(f"{1e3*t:.2f}" for t in ts) + f" ms  pairs {n} area {area:.12f}  best {n/min(ts)/1e6:.1f} Mpairs/s")
