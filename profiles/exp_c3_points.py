"""Experiment: locate_points on the 2 M-triangle Delaunay tree (C3), 50 M points, device-resident."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import delaunay_mesh
n = int(os.environ.get("NPTS", 50_000_000))
v, f = delaunay_mesh(1_000_000, 1234)
tree = CellTree2d(v, f, -1)
pts = torch.from_numpy(np.random.default_rng(7).uniform(0, 1, (n, 2))).cuda()
for _ in range(3): out = tree.locate_points(pts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = tree.locate_points(pts)
e1.record(); torch.cuda.synchronize()
lib = _lib.load(); lib.ct_profile_enable(1); tree.locate_points(pts)
a, b = ctypes.c_double(), ctypes.c_double(); lib.ct_profile_last(ctypes.byref(a), ctypes.byref(b)); lib.ct_profile_enable(0)
print("C3 locate_points ms/step %.3f" % (e0.elapsed_time(e1) / 5), "Gq/s %.3f" % (n / (e0.elapsed_time(e1) / 5) / 1e6),
      "order %.3f traverse %.3f" % (a.value, b.value), "found", int((out >= 0).sum().item()))
