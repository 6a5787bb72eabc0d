"""Experiment: tree build timing (cold / warm pool) on the C2 and C3 meshes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CELLTREE_DEBUG"] = "1"
import numpy as np
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import quad_mesh, delaunay_mesh
nx = int(os.environ.get("NX", 4096))
v, f = quad_mesh(nx, nx)
for k in range(3):
    t0 = time.perf_counter(); t = CellTree2d(v, f, -1); dt = time.perf_counter() - t0
    print(f"quad {nx}: total {dt*1e3:.1f} ms, device build {t.build_ms:.1f} ms, nodes {t._tree.info.n_nodes}, depth {t.depth}", flush=True)
    del t
v, f = delaunay_mesh(int(os.environ.get("NPTS", 1_000_000)), 1234)
for k in range(3):
    t0 = time.perf_counter(); t = CellTree2d(v, f, -1); dt = time.perf_counter() - t0
    print(f"delaunay: total {dt*1e3:.1f} ms, device build {t.build_ms:.1f} ms, nodes {t._tree.info.n_nodes}, depth {t.depth}", flush=True)
    del t
