"""Small run of every query type for compute-sanitizer (no torch: NumPy in, NumPy out), checked against the oracle.

    compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log python profiles/sanitize_small.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from numba_celltree_b200 import CellTree2d, EdgeCellTree2d
from numba_celltree_b200.synthetic import c3_boxes, c4_edges, delaunay_mesh, quad_mesh, random_network

t0 = time.perf_counter()
def stage(name):
    print(f"{time.perf_counter() - t0:6.1f} s  {name}", flush=True)

vertices, faces = delaunay_mesh(1500, seed=3)
tree, ref = CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1)
assert np.array_equal(tree.bb_indices, ref.bb_indices)
stage("build")
edges = c4_edges(len(faces), 3000)
got, want = tree.intersect_edges(edges), ref.intersect_edges(edges)
assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, want))
stage(f"intersect_edges ({len(got[0])} pairs)")
points = np.random.default_rng(0).uniform(-0.05, 1.05, (6000, 2))
got, want = tree.compute_barycentric_weights(points), ref.compute_barycentric_weights(points)
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
stage("locate_points + weights")
boxes = c3_boxes(len(faces), 3000)
got, want = tree.intersect_boxes(boxes), ref.intersect_boxes(boxes)
assert all(np.array_equal(a, b) for a, b in zip(got, want))
stage(f"intersect_boxes ({len(got[0])} pairs)")
qv, qf = quad_mesh(25, 20)
got, want = tree.intersect_faces(qv, qf, -1), ref.intersect_faces(qv, qf, -1)
assert all(np.array_equal(a, b) for a, b in zip(got, want))
stage(f"intersect_faces ({len(got[0])} pairs)")
nv, ne = random_network(600, seed=2)
net, net_ref = EdgeCellTree2d(nv, ne), oracle.EdgeCellTree2d(nv, ne)
a = np.random.default_rng(1).uniform(nv.min(0), nv.max(0), (2000, 2))
segs = np.stack((a, a + np.random.default_rng(2).normal(0, 3.0, a.shape)), axis=1)
got, want = net.intersect_edges(segs), net_ref.intersect_edges(segs)
assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(got, want))
assert np.array_equal(net.locate_points(a), net_ref.locate_points(a))
stage(f"network queries ({len(got[0])} pairs)")
print("sanitize_small ok")
