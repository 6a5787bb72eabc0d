"""Experiment: per-kernel timing of locate_points on C2 under different knobs (run under ncu for the launch list)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import quad_mesh
nx = int(os.environ.get("NX", 4096)); n = int(os.environ.get("NPTS", 100_000_000))
v, f = quad_mesh(nx, nx)
tree = CellTree2d(v, f, -1)
pts = torch.from_numpy(np.random.default_rng(42).uniform(0, 1, (n, 2))).cuda()
for _ in range(3): out = tree.locate_points(pts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = tree.locate_points(pts)
e1.record(); torch.cuda.synchronize()
print("ms/step", e0.elapsed_time(e1) / 5, "Gq/s", n / (e0.elapsed_time(e1) / 5) / 1e6)
