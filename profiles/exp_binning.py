"""Experiment: device time of the point binning alone (CELLTREE_B200_LIB selects a library variant)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import quad_mesh
nx = int(os.environ.get("NX", 1024)); n = int(os.environ.get("NPTS", 100_000_000))
v, f = quad_mesh(nx, nx)
tree = CellTree2d(v, f, -1)
lib = _lib.load()
lib.ct_profile_binning.restype = ctypes.c_int
lib.ct_profile_binning.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for name in os.environ.get("INPUTS", "uniform,sorted,constant").split(","):
    pts = torch.rand((n, 2), dtype=torch.float64, device="cuda", generator=gen)
    if name == "sorted":
        pts = pts[torch.argsort((pts[:, 1] * 4096).floor() * 4096 + (pts[:, 0] * 4096).floor())].contiguous()
    if name == "constant":
        pts[:] = 0.3
    ms = ctypes.c_double()
    _lib.check(lib.ct_profile_binning(tree._tree.handle, pts.data_ptr(), n, 5, ctypes.byref(ms)))
    print(os.environ.get("CELLTREE_B200_LIB", "default").split("/")[-1], name, "binning ms %.3f" % ms.value)
    del pts
