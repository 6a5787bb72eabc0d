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
# round 2: the spatial binning of large point batches (forced on for this small batch), every execution order, results
# through the window queues (direct-out limit 0) and directly; crowded points overflow their slabs; a deep tree walks on
# the overflow slab of the traversal stacks; signed-zero build; the batched geometry helpers; an edited node array
from numba_celltree_b200 import _lib, algorithms
lib = _lib.load()
lib.ct_set_sort_bits(16)
crowded = np.concatenate([points, np.random.default_rng(5).normal(0.4, 0.0005, (5000, 2))])
want = ref.compute_barycentric_weights(crowded)
got = tree.compute_barycentric_weights(crowded)
assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
assert np.array_equal(net.locate_points(a), net_ref.locate_points(a))
lib.ct_set_sort_bits(-1)
stage("binned points (slabs, overflow, weights, edge tree)")
x = 1.0 / 2.0 ** np.arange(150)
k = np.arange(150)
dv = np.concatenate([np.column_stack((-x, np.zeros(150))), np.column_stack((3 * x, np.zeros(150))), np.column_stack((3 * x, np.ones(150))), np.column_stack((-x, np.ones(150)))])
df = np.column_stack((k, k + 150, k + 300, k + 450))
deep, deep_ref = CellTree2d(dv, df, -1), oracle.CellTree2d(dv, df, -1)
rng = np.random.default_rng(9)
dp = np.column_stack((rng.choice([-1.0, 1.0], 3000) * 2.0 ** -rng.uniform(0, 155, 3000), rng.uniform(0, 1, 3000)))
assert deep.depth > 64 and np.array_equal(deep.locate_points(dp), deep_ref.locate_points(dp))
db = np.column_stack((-(2.0 ** -rng.uniform(0, 150, 300)), 2.0 ** -rng.uniform(0, 150, 300), rng.uniform(0, 0.4, 300), rng.uniform(0.5, 1, 300)))
got, want = deep.intersect_boxes(db), deep_ref.intersect_boxes(db)
assert all(np.array_equal(p_, q_) for p_, q_ in zip(got, want))
ds = np.stack((np.column_stack((-(2.0 ** -rng.uniform(0, 150, 200)), rng.uniform(0, 1, 200))), np.column_stack((2.0 ** -rng.uniform(0, 150, 200), rng.uniform(0, 1, 200)))), axis=1)
got, want = deep.intersect_edges(ds), deep_ref.intersect_edges(ds)
assert all(np.array_equal(p_, q_, equal_nan=True) for p_, q_ in zip(got, want))
stage(f"deep tree ({deep.depth} levels)")
zv, zf = quad_mesh(8, 6)
zv = zv * [8, 6] - [4, 3]
zv[(zv[:, 0] == 0) & (np.arange(len(zv)) % 2 == 0), 0] = -0.0
zt, zr = CellTree2d(zv, zf, -1, cells_per_leaf=1), oracle.CellTree2d(zv, zf, -1, cells_per_leaf=1)
assert zt.nodes.tobytes() == zr.nodes.tobytes()
inner = int(np.flatnonzero(ref.nodes["child"] != -1)[2])
tree.nodes["Lmax"][inner] = -1.0
ref.nodes["Lmax"][inner] = -1.0
assert np.array_equal(tree.locate_points(points), ref.locate_points(points))
hit, c, d = algorithms.liang_barsky_line_box_clip(edges[:, 0], edges[:, 1], boxes[: len(edges)])
want = oracle.liang_barsky_line_box_clip(edges[:, 0], edges[:, 1], boxes[: len(edges)])
assert np.array_equal(hit, want[0]) and np.array_equal(c, want[1], equal_nan=True)
assert np.array_equal(algorithms.points_in_polygon(points, vertices[faces[0]]), oracle.points_in_polygon(points, vertices[faces[0]]))
stage("signed zeros, node edit, geometry helpers")
print("sanitize_small ok")
