"""Experiment: full-size C3 / C4 / C5 (boxes, edges, faces on the 2M-triangle Delaunay tree), device-resident inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numba_celltree_b200 import CellTree2d
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh, c3_boxes, c4_edges
npts = int(os.environ.get("NPTS", 1_000_000)); nq = int(os.environ.get("NQ", 10_000_000))
t0 = time.perf_counter(); v, f = delaunay_mesh(npts, 1234); print("delaunay s", time.perf_counter() - t0, len(f), flush=True)
tree = CellTree2d(v, f, -1); print("build ms", tree.build_ms, "depth", tree.depth, flush=True)
def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r
boxes = torch.from_numpy(c3_boxes(len(f), nq)).cuda()
t, r = timeit("locate_boxes", lambda: tree.locate_boxes(boxes)); print(f"locate_boxes    {t*1e3:9.2f} ms  {nq/t/1e6:8.2f} Mbox/s  {len(r[0])/t/1e6:8.1f} Mpairs/s  pairs {len(r[0])}", flush=True)
t, r = timeit("intersect_boxes", lambda: tree.intersect_boxes(boxes)); print(f"intersect_boxes {t*1e3:9.2f} ms  {nq/t/1e6:8.2f} Mbox/s  {len(r[0])/t/1e6:8.1f} Mpairs/s  pairs {len(r[0])}", flush=True)
del boxes, r
edges = torch.from_numpy(c4_edges(len(f), nq)).cuda()
t, r = timeit("intersect_edges", lambda: tree.intersect_edges(edges)); print(f"intersect_edges {t*1e3:9.2f} ms  {nq/t/1e6:8.2f} Medge/s {len(r[0])/t/1e6:8.1f} Mpairs/s  pairs {len(r[0])}", flush=True)
del edges, r
qv, qf = quad_mesh(1000, 1000)
t, r = timeit("intersect_faces", lambda: tree.intersect_faces(qv, qf, -1)); print(f"intersect_faces {t*1e3:9.2f} ms  {len(qf)/t/1e6:8.2f} Mface/s {len(r[0])/t/1e6:8.1f} Mpairs/s  pairs {len(r[0])} area {r[2].sum():.12f}", flush=True)
pts = torch.from_numpy(np.random.default_rng(7).uniform(0, 1, (nq * 5, 2))).cuda()
t, r = timeit("locate_points", lambda: tree.locate_points(pts)); print(f"locate_points   {t*1e3:9.2f} ms  {nq*5/t/1e6:8.2f} Mq/s", flush=True)
t, r = timeit("weights", lambda: tree.compute_barycentric_weights(pts)); print(f"bary weights    {t*1e3:9.2f} ms  {nq*5/t/1e6:8.2f} Mq/s", flush=True)
