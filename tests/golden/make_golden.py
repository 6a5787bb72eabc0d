"""
Generate the golden fixtures in tests/golden/*.npz by running the REFERENCE
(Deltares/numba_celltree v0.4.2, mounted read-only at /root/reference) under
Numba in the build container.  The GPU box has no /root/reference: only the
.npz files travel.

    NUMBA_CACHE_DIR=/tmp/nbcache python tests/golden/make_golden.py

Every fixture holds the seeded inputs and the reference's outputs for one mesh:
tree arrays (nodes / bb_indices / bb_coords / faces after counter_clockwise)
for several (n_buckets, cells_per_leaf), and the results of every query type in
the reference's output order.  Meshes are either small analytic recipes, seeded
random point sets, or the two small real meshes the reference's own tests use
(tests/data: 538 triangles, 74 Voronoi polygons) stored here as input vectors.
"""

import os
import pathlib
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
REFERENCE = os.environ.get("CELLTREE_REFERENCE", "/root/reference")
sys.path.insert(0, REFERENCE)

import numpy as np  # noqa: E402
import scipy.spatial  # noqa: E402
from numba_celltree import CellTree2d, EdgeCellTree2d  # noqa: E402

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from numba_celltree_b200.synthetic import delaunay_mesh, generate_disk, quad_mesh, random_network  # noqa: E402

BUILD_PARAMS = [(4, 2), (2, 1), (2, 2), (8, 3), (3, 1), (16, 4)]


def tree_arrays(tree, prefix, out):
    out[f"{prefix}nodes"] = tree.nodes
    out[f"{prefix}bb_indices"] = tree.bb_indices
    out[f"{prefix}bb_coords"] = tree.bb_coords
    out[f"{prefix}bbox"] = tree.bbox


def face_tree_case(name, vertices, faces, fill_value, rng, n_query=400, other=None, extra_points=None):
    """One fixture: CellTree2d on (vertices, faces) + every query type."""
    out = {"vertices": vertices, "faces": faces, "fill_value": np.int64(fill_value)}
    for nb_, cpl in BUILD_PARAMS:
        t = CellTree2d(vertices, faces, fill_value, n_buckets=nb_, cells_per_leaf=cpl)
        tree_arrays(t, f"b{nb_}_c{cpl}_", out)
    tree = CellTree2d(vertices, faces, fill_value)
    out["faces_ccw"] = tree.faces
    out["bb_distances"] = tree.bb_distances
    xmin, xmax, ymin, ymax = tree.bbox
    dx, dy = xmax - xmin, ymax - ymin

    # points: uniform over a slightly larger box, every vertex, every edge midpoint, every centroid
    pts = rng.uniform((xmin - 0.1 * dx, ymin - 0.1 * dy), (xmax + 0.1 * dx, ymax + 0.1 * dy), (n_query * 5, 2))
    f = tree.faces
    closed = np.where(f == -1, f[:, :1], f)
    mid = 0.5 * (vertices[closed] + vertices[np.roll(closed, -1, axis=1)]).reshape(-1, 2)
    cen = np.array([vertices[row[row != -1]].mean(axis=0) for row in f])
    points = np.concatenate([pts, vertices, mid[:: max(1, len(mid) // 2000)], cen[:: max(1, len(cen) // 2000)]])
    if extra_points is not None:
        points = np.concatenate([points, extra_points])
    out["points"] = points
    out["locate_points"] = tree.locate_points(points)
    for k, tol in enumerate((1e-9, 1e-3 * max(dx, dy))):
        out[f"locate_points_tol{k}"] = tree.locate_points(points, tolerance=tol)
        out[f"tol{k}"] = np.float64(tol)
    fi, w = tree.compute_barycentric_weights(points)
    out["weights"] = w
    assert np.array_equal(fi, out["locate_points"])

    # boxes
    h = np.sqrt(dx * dy / len(faces))
    c = rng.uniform((xmin - 0.05 * dx, ymin - 0.05 * dy), (xmax + 0.05 * dx, ymax + 0.05 * dy), (n_query, 2))
    wh = rng.uniform(0, 6 * h, (n_query, 2))
    boxes = np.column_stack((c[:, 0] - wh[:, 0] / 2, c[:, 0] + wh[:, 0] / 2, c[:, 1] - wh[:, 1] / 2, c[:, 1] + wh[:, 1] / 2))
    # a few boxes exactly aligned to cell bounding boxes (touching is not intersecting) and one covering everything
    boxes = np.concatenate([boxes, tree.bb_coords[:: max(1, len(faces) // 50)], [[xmin - 1, xmax + 1, ymin - 1, ymax + 1]]])
    out["boxes"] = boxes
    out["locate_boxes_i"], out["locate_boxes_j"] = tree.locate_boxes(boxes)
    out["intersect_boxes_i"], out["intersect_boxes_j"], out["intersect_boxes_area"] = tree.intersect_boxes(boxes)

    # edges: random segments, plus segments through vertices, along cell edges, degenerate
    a0 = rng.uniform((xmin - 0.05 * dx, ymin - 0.05 * dy), (xmax + 0.05 * dx, ymax + 0.05 * dy), (n_query, 2))
    ang = rng.uniform(0, 2 * np.pi, n_query)
    L = rng.uniform(0, 12 * h, n_query)
    edges = np.stack((a0, a0 + np.column_stack((np.cos(ang), np.sin(ang))) * L[:, None]), axis=1)
    step = max(1, len(faces) // 40)
    along = np.stack((vertices[closed[::step, 0]], vertices[closed[::step, 1]]), axis=1)  # collinear with a cell edge
    through = np.stack((cen[::step], vertices[closed[::step, 0]]), axis=1)  # centroid -> vertex
    through2 = np.stack((2 * cen[::step] - vertices[closed[::step, 0]], 2 * vertices[closed[::step, 0]] - cen[::step]), axis=1)
    axis_h = np.array([[[xmin - 1, ymin + 0.37 * dy], [xmax + 1, ymin + 0.37 * dy]], [[xmin + 0.5 * dx, ymax + 1], [xmin + 0.5 * dx, ymin - 1]]])
    degenerate = np.array([[[xmin + 0.5 * dx, ymin + 0.5 * dy]] * 2])
    edges = np.concatenate([edges, along, along[:, ::-1], through, through2, axis_h, degenerate])
    out["edges"] = edges
    out["intersect_edges_i"], out["intersect_edges_j"], out["intersect_edges_xy"] = tree.intersect_edges(edges)

    # faces: another mesh laid over this one
    if other is None:
        nq = 24
        qv, qf = quad_mesh(nq, nq)
        qv = qv * [1.1 * dx, 1.1 * dy] + [xmin - 0.05 * dx, ymin - 0.05 * dy]
        other = (qv, qf, -1)
    ov, of, ofill = other
    out["other_vertices"], out["other_faces"], out["other_fill"] = ov, of, np.int64(ofill)
    out["intersect_faces_i"], out["intersect_faces_j"], out["intersect_faces_area"] = tree.intersect_faces(ov, of, ofill)
    of_c = np.where(of == ofill, -1, of).astype(np.intp)
    li, lj = tree.locate_faces(ov.copy(), of_c)
    out["locate_faces_i"], out["locate_faces_j"] = li, lj
    # the mesh against itself: pair membership hangs on rounding noise
    si, sj, sa = tree.intersect_faces(vertices, faces, fill_value)
    out["self_faces_i"], out["self_faces_j"], out["self_faces_area"] = si, sj, sa

    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0 and not k.startswith("b")})


def edge_tree_case(name, vertices, edges, rng, n_query=400):
    out = {"vertices": vertices, "edges": edges}
    for nb_, cpl in BUILD_PARAMS:
        t = EdgeCellTree2d(vertices, edges, n_buckets=nb_, cells_per_leaf=cpl)
        tree_arrays(t, f"b{nb_}_c{cpl}_", out)
    tree = EdgeCellTree2d(vertices, edges)
    out["bb_distances"] = tree.bb_distances
    xmin, xmax, ymin, ymax = tree.bbox
    dx, dy = xmax - xmin, ymax - ymin
    p = vertices[edges[:, 0]]
    q = vertices[edges[:, 1]]
    s = rng.uniform(0, 1, (len(edges), 1))
    on = p + s * (q - p)
    off = on + rng.normal(0, 1e-3 * max(dx, dy), on.shape)
    rnd = rng.uniform((xmin, ymin), (xmax, ymax), (n_query, 2))
    points = np.concatenate([on, off, rnd, vertices])
    out["points"] = points
    out["locate_points"] = tree.locate_points(points)
    for k, tol in enumerate((1e-9, 2e-3 * max(dx, dy))):
        out[f"locate_points_tol{k}"] = tree.locate_points(points, tolerance=tol)
        out[f"tol{k}"] = np.float64(tol)
    a0 = rng.uniform((xmin, ymin), (xmax, ymax), (n_query, 2))
    ang = rng.uniform(0, 2 * np.pi, n_query)
    L = rng.uniform(0, 0.3 * max(dx, dy), n_query)
    qe = np.stack((a0, a0 + np.column_stack((np.cos(ang), np.sin(ang))) * L[:, None]), axis=1)
    k = max(1, len(edges) // 40)
    same = np.stack((p[::k], q[::k]), axis=1)  # identical to tree edges (collinear)
    part = np.stack((p[::k] + 0.25 * (q[::k] - p[::k]), p[::k] + 1.5 * (q[::k] - p[::k])), axis=1)  # collinear, partial overlap
    touch = np.stack((p[::k], p[::k] + (q[::k] - p[::k])[:, ::-1] * [1, -1]), axis=1)  # starts on a vertex, orthogonal
    degenerate = np.array([[[xmin + 0.5 * dx, ymin + 0.5 * dy]] * 2])
    qe = np.concatenate([qe, same, same[:, ::-1], part, touch, degenerate])
    out["query_edges"] = qe
    out["intersect_edges_i"], out["intersect_edges_j"], out["intersect_edges_xy"] = tree.intersect_edges(qe)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0 and not k.startswith("b")})


def main():
    data = pathlib.Path(REFERENCE) / "tests" / "data"
    tri_xy = np.loadtxt(data / "xy.txt", dtype=float)
    tri = np.loadtxt(data / "triangles.txt", dtype=int)
    vor_xy = np.loadtxt(data / "voronoi_xy.txt", dtype=float)
    vor = np.loadtxt(data / "voronoi.txt", dtype=int)

    rng = np.random.default_rng(20261017)

    # 1. the reference's example mesh: generate_disk(5, 5), re-sorted by centroid distance, scaled
    v, f = generate_disk(5, 5)
    centroids = v[f].mean(axis=1)
    order = np.argsort((centroids[:, 0] + 1.0) ** 2 + (centroids[:, 1] + 1.0) ** 2)
    f = f[order]
    v = (v + 1.0) * 5.0
    face_tree_case("disk_5_5", v, f, -1, rng)

    # 2./3. the reference's two real test meshes, each laid over the other
    face_tree_case("triangles_538", tri_xy, tri, -1, rng, other=(vor_xy, vor, -1))
    face_tree_case("voronoi_74", vor_xy, vor, -1, rng, other=(tri_xy, tri, -1))

    # 4. structured quads, points exactly on edges and vertices, -999 fill mixed with triangles
    qv, qf = quad_mesh(48, 40)
    qf = np.column_stack([qf, np.full(len(qf), -999)])
    tri_rows = rng.choice(len(qf), 200, replace=False)
    qf[tri_rows, 3] = -999  # degrade some quads to triangles: mixed arity
    face_tree_case("quads_48_40_mixed", qv * [3.0, 2.0] + [10.0, -5.0], qf, -999, rng)

    # 5. seeded random Delaunay (overlapping children, leaves of 1..4 cells), vs itself
    dv, df = delaunay_mesh(3000, seed=7)
    face_tree_case("delaunay_3000", dv, df, -1, rng, n_query=1500)

    # 6. duplicate cells: identical centroids force the retry / oversized-leaf branches
    base_v, base_f = delaunay_mesh(60, seed=3)
    dup_f = np.concatenate([base_f, base_f[:40], base_f[:40], base_f[:17][:, [1, 2, 0]], base_f[:5]])
    face_tree_case("duplicates", base_v, dup_f, -1, rng, n_query=200)

    # 7. edge networks
    from_demo_v = np.array(
        [[0.0, 0.0], [0.25, 1.0], [1.25, 2.0], [1.5, 2.5], [2.5, 3.25], [2.5, 2.5], [2.75, 3.75], [3.0, 2.0], [0.25, 1.75], [0.5, 2.25]]
    )
    from_demo_e = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [3, 5], [4, 6], [5, 7], [1, 8], [8, 9]], dtype=np.int64)
    edge_tree_case("network_demo", from_demo_v, from_demo_e, rng, n_query=100)
    nv, ne = random_network(800, seed=11)
    edge_tree_case("network_800", nv, ne, rng)
    # axis-aligned network (zero-width boxes padded by the global tolerance)
    gv, gf = quad_mesh(12, 9)
    ge = np.unique(np.sort(np.concatenate([gf[:, [0, 1]], gf[:, [1, 2]], gf[:, [2, 3]], gf[:, [3, 0]]]), axis=1), axis=0)
    edge_tree_case("network_grid", gv * 100.0 + 5000.0, ge.astype(np.int64), rng, n_query=200)


if __name__ == "__main__" and os.environ.get("CELLTREE_GOLDEN_ONLY", "all") == "all":
    main()


# ---------------------------------------------------------------------------------------------------------
# Degenerate lattice cases: every coordinate on a (half-)integer lattice, so that queries run through
# vertices, along cell edges, touch boxes exactly and overlap collinearly -- the branches that random
# inputs never reach.  Expected outputs come from the reference, as everywhere else in this file.
def lattice_cases():
    # a 4 x 3 block of unit cells: quads, some split into two triangles, one pentagon with a hanging node
    xs, ys = np.meshgrid(np.arange(5.0), np.arange(4.0), indexing="xy")
    vertices = np.column_stack((xs.ravel(), ys.ravel()))
    vertices = np.vstack([vertices, [[2.5, 0.0]]])  # vertex 20: hanging node on the bottom edge of cell (2, 0)
    vid = lambda i, j: j * 5 + i  # noqa: E731
    faces = []
    for j in range(3):
        for i in range(4):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            if (i, j) == (2, 0):
                faces.append([a, 20, b, c, d])  # pentagon with a collinear (hanging) vertex
            elif (i + j) % 3 == 0:
                faces.append([a, b, c, -1, -1])
                faces.append([a, c, d, -1, -1])
            elif (i + j) % 3 == 1:
                faces.append([d, c, b, a, -1])  # clockwise quad: counter_clockwise must flip it
            else:
                faces.append([a, b, c, d, -1])
    faces = np.array(faces, dtype=np.int64)
    out = {"vertices": vertices, "faces": faces, "fill_value": np.int64(-1)}
    for nb_, cpl in BUILD_PARAMS:
        t = CellTree2d(vertices, faces, -1, n_buckets=nb_, cells_per_leaf=cpl)
        tree_arrays(t, f"b{nb_}_c{cpl}_", out)
    tree = CellTree2d(vertices, faces, -1)
    out["faces_ccw"] = tree.faces
    out["bb_distances"] = tree.bb_distances
    hx, hy = np.meshgrid(np.arange(-1.0, 5.01, 0.5), np.arange(-1.0, 4.01, 0.5), indexing="xy")
    points = np.column_stack((hx.ravel(), hy.ravel()))
    eps = np.array([[1e-13, 0.0], [0.0, -1e-13], [3e-16, 3e-16], [-1e-9, 1e-9]])
    points = np.concatenate([points] + [points + e for e in eps])
    out["points"] = points
    out["locate_points"] = tree.locate_points(points)
    for k, tol in enumerate((1e-9, 1e-3)):
        out[f"locate_points_tol{k}"] = tree.locate_points(points, tolerance=tol)
        out[f"tol{k}"] = np.float64(tol)
    _, out["weights"] = tree.compute_barycentric_weights(points)
    _, out["weights_tol1"] = tree.compute_barycentric_weights(points, tolerance=1e-3)
    c = np.arange(-1.0, 5.01, 1.0)
    boxes = np.array([[x0, x1, y0, y1] for x0 in c for x1 in c if x1 > x0 for y0 in c[:-1] for y1 in c[:-1] if y1 > y0])
    boxes = np.concatenate([boxes, boxes[::7] + [0.5, 0.5, 0.5, 0.5], [[1.0, 1.0, 0.0, 2.0]]])
    out["boxes"] = boxes
    out["locate_boxes_i"], out["locate_boxes_j"] = tree.locate_boxes(boxes)
    out["intersect_boxes_i"], out["intersect_boxes_j"], out["intersect_boxes_area"] = tree.intersect_boxes(boxes)
    px, py = np.meshgrid(np.arange(-0.5, 4.51, 0.5), np.arange(-0.5, 3.51, 0.5), indexing="xy")
    lp = np.column_stack((px.ravel(), py.ravel()))
    ia, ib = np.meshgrid(np.arange(len(lp)), np.arange(len(lp)), indexing="ij")
    sel = (ia.ravel() * 7 + ib.ravel() * 3) % 5 == 0  # every fifth ordered pair, degenerate ones included
    edges = np.stack((lp[ia.ravel()[sel]], lp[ib.ravel()[sel]]), axis=1)
    out["edges"] = edges
    out["intersect_edges_i"], out["intersect_edges_j"], out["intersect_edges_xy"] = tree.intersect_edges(edges)
    # the same mesh shifted by (half-)integer offsets: shared edges, shared vertices, exact containment
    ov = np.concatenate([vertices + s for s in ([0.0, 0.0], [0.5, 0.0], [1.0, 1.0], [0.5, 0.5], [4.0, 0.0], [-1.0, 3.0])])
    of = np.concatenate([np.where(faces == -1, -1, faces + k * len(vertices)) for k in range(6)])
    out["other_vertices"], out["other_faces"], out["other_fill"] = ov, of, np.int64(-1)
    out["intersect_faces_i"], out["intersect_faces_j"], out["intersect_faces_area"] = tree.intersect_faces(ov, of, -1)
    li, lj = tree.locate_faces(ov.copy(), of.copy())
    out["locate_faces_i"], out["locate_faces_j"] = li, lj
    si, sj, sa = tree.intersect_faces(vertices, faces, -1)
    out["self_faces_i"], out["self_faces_j"], out["self_faces_area"] = si, sj, sa
    np.savez_compressed(HERE / "lattice_faces.npz", **out)
    print("lattice_faces", len(points), len(boxes), len(edges), len(out["intersect_edges_i"]), len(out["intersect_faces_i"]))

    # lattice network for EdgeCellTree2d: unit horizontal / vertical / diagonal edges, one zero-length edge
    nv = np.column_stack((xs.ravel(), ys.ravel()))
    ne = []
    for j in range(4):
        for i in range(5):
            if i < 4 and (i + j) % 2 == 0:
                ne.append([vid(i, j), vid(i + 1, j)])
            if j < 3 and (i + 2 * j) % 3 != 1:
                ne.append([vid(i, j), vid(i, j + 1)])
            if i < 4 and j < 3 and (i * j) % 2 == 1:
                ne.append([vid(i, j), vid(i + 1, j + 1)])
    ne.append([7, 7])
    ne = np.array(ne, dtype=np.int64)
    out = {"vertices": nv, "edges": ne}
    for nb_, cpl in BUILD_PARAMS:
        t = EdgeCellTree2d(nv, ne, n_buckets=nb_, cells_per_leaf=cpl)
        tree_arrays(t, f"b{nb_}_c{cpl}_", out)
    tree = EdgeCellTree2d(nv, ne)
    out["bb_distances"] = tree.bb_distances
    out["points"] = points
    out["locate_points"] = tree.locate_points(points)
    for k, tol in enumerate((1e-9, 1e-3)):
        out[f"locate_points_tol{k}"] = tree.locate_points(points, tolerance=tol)
        out[f"tol{k}"] = np.float64(tol)
    out["query_edges"] = edges
    out["intersect_edges_i"], out["intersect_edges_j"], out["intersect_edges_xy"] = tree.intersect_edges(edges)
    np.savez_compressed(HERE / "lattice_network.npz", **out)
    print("lattice_network", len(ne), len(out["intersect_edges_i"]))


if __name__ == "__main__" and os.environ.get("CELLTREE_GOLDEN_ONLY", "all") in ("all", "lattice"):
    lattice_cases()
