"""Experiment: locate_points on C2, step / ordering / traversal time (CELLTREE_B200_LIB selects a library variant)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import quad_mesh
nx = int(os.environ.get("NX", 4096)); n = int(os.environ.get("NPTS", 100_000_000))
v, f = quad_mesh(nx, nx)
if os.environ.get("L2FETCH"):
    import glob
    path = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))[0]
    rt = ctypes.CDLL(path)
    torch.cuda.init(); torch.zeros(1).cuda()
    got = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2FETCH"])))
    now = ctypes.c_size_t(); rt.cudaDeviceGetLimit(ctypes.byref(now), 5)
    print("cudaLimitMaxL2FetchGranularity", got.value, "->", now.value, "rc", rc)
tree = CellTree2d(v, f, -1)
tree2 = CellTree2d(v, f, -1)
pts = torch.from_numpy(np.random.default_rng(42).uniform(0, 1, (n, 2))).cuda()
for _ in range(3): out = tree.locate_points(pts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = tree.locate_points(pts)
e1.record(); torch.cuda.synchronize()
lib = _lib.load()
lib.ct_profile_enable(1)
tree.locate_points(pts)
a, b = ctypes.c_double(), ctypes.c_double()
lib.ct_profile_last(ctypes.byref(a), ctypes.byref(b))
lib.ct_profile_enable(0)
print(os.environ.get("CELLTREE_B200_LIB", "default").split("/")[-1], "ms/step %.3f" % (e0.elapsed_time(e1) / 5),
      "Gq/s %.3f" % (n / (e0.elapsed_time(e1) / 5) / 1e6), "order %.3f traverse %.3f" % (a.value, b.value),
      "build %.1f ms" % tree2.build_ms, "checksum", int(out.sum().item()))

if os.environ.get("WEIGHTS", "1") != "0":
    for _ in range(2): r = tree.compute_barycentric_weights(pts)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3): r = tree.compute_barycentric_weights(pts)
    e1.record(); torch.cuda.synchronize()
    print("compute_barycentric_weights ms/step %.3f" % (e0.elapsed_time(e1) / 3), "Gq/s %.3f" % (n / (e0.elapsed_time(e1) / 3) / 1e6),
          "weights checksum %.9f" % float(r[1].sum().item()))
