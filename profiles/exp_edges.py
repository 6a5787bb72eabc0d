"""Experiment: intersect_edges on C4 (subset), for ncu captures and stage timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import delaunay_mesh, c4_edges
nq = int(os.environ.get("NQ", 2_000_000))
v, f = delaunay_mesh(1_000_000, 1234)
tree = CellTree2d(v, f, -1)
edges = torch.from_numpy(c4_edges(len(f), nq)).cuda()
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    i, j, xy = tree.intersect_edges(edges)
    torch.cuda.synchronize(); print(f"--- intersect_edges {1e3*(time.perf_counter()-t0):.2f} ms pairs {len(i)} ({nq/(time.perf_counter()-t0)/1e6:.1f} Medge/s)", file=sys.stderr, flush=True)
