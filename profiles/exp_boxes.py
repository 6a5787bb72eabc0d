"""Experiment: stage timing of intersect_boxes on C3 (CELLTREE_DEBUG trace points)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import delaunay_mesh, c3_boxes
v, f = delaunay_mesh(1_000_000, 1234)
tree = CellTree2d(v, f, -1)
boxes = torch.from_numpy(c3_boxes(len(f), 10_000_000)).cuda()
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    i, j, a = tree.intersect_boxes(boxes)
    torch.cuda.synchronize(); print(f"--- intersect_boxes {1e3*(time.perf_counter()-t0):.2f} ms pairs {len(i)}", file=sys.stderr, flush=True)
    del i, j, a
for k in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    i, j = tree.locate_boxes(boxes)
    torch.cuda.synchronize(); print(f"--- locate_boxes {1e3*(time.perf_counter()-t0):.2f} ms pairs {len(i)}", file=sys.stderr, flush=True)
