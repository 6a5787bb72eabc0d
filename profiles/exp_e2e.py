"""Experiment: end-to-end locate_points with pinned host buffers (C2), and the raw PCIe copy rates beside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import quad_mesh
n = int(os.environ.get("NPTS", 100_000_000))
v, f = quad_mesh(4096, 4096)
tree = CellTree2d(v, f, -1)
hp = torch.empty((n, 2), dtype=torch.float64).pin_memory()
np.random.default_rng(42).random(out=hp.numpy().reshape(-1))
ho = torch.empty(n, dtype=torch.int64).pin_memory()
dev = torch.empty_like(hp, device="cuda"); dout = torch.empty(n, dtype=torch.int64, device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts)
a = t(lambda: dev.copy_(hp, non_blocking=True)); b = t(lambda: ho.copy_(dout, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dev.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(dout, non_blocking=True)
c = t(both)
print(f"H2D 1.6 GB {a*1e3:.2f} ms = {1.6*n/1e8/a:.1f} GB/s; D2H 0.8 GB {b*1e3:.2f} ms = {0.8*n/1e8/b:.1f} GB/s; both at once {c*1e3:.2f} ms")
e = t(lambda: tree.locate_points(hp.numpy(), out=ho.numpy()), reps=5)
print(f"CELLTREE_HOST_CHUNK={os.environ.get('CELLTREE_HOST_CHUNK','default')}: e2e {e*1e3:.2f} ms = {n/e/1e9:.3f} Gq/s")
