"""
CUDA path vs the CPU oracle on seeded random inputs at sizes the oracle finishes in seconds, plus the edge
cases the domain has: empty query sets, a single cell, everything outside, NaN coordinates, trees deeper than
the per-thread traversal stack, device-resident inputs, and Morton ordering on/off.
"""

import numpy as np
import pytest

import oracle
from numba_celltree_b200.synthetic import c3_boxes, c4_edges, delaunay_mesh, quad_mesh
from tests._golden import RTOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import numba_celltree_b200

    return numba_celltree_b200


@pytest.fixture(scope="module")
def delaunay_pair(pkg):
    vertices, faces = delaunay_mesh(100_000, seed=1234)
    return pkg.CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1), vertices, faces


def assert_same_tree(tree, ref):
    for field in ("child", "Lmax", "Rmin", "ptr", "size", "dim"):
        assert np.array_equal(tree.nodes[field], ref.nodes[field]), field
    assert np.array_equal(tree.bb_indices, ref.bb_indices)
    assert np.array_equal(tree.bb_coords, ref.bb_coords)
    assert np.array_equal(tree.bbox, ref.bbox)


def test_build_delaunay_bit_exact(delaunay_pair):
    tree, ref, _, _ = delaunay_pair
    assert_same_tree(tree, ref)
    assert np.array_equal(tree.faces, ref.faces)


@pytest.mark.parametrize("n_buckets,cells_per_leaf", [(2, 1), (3, 2), (8, 4), (13, 1), (64, 2), (100, 3)])
def test_build_parameters_bit_exact(pkg, n_buckets, cells_per_leaf):
    vertices, faces = delaunay_mesh(20_000, seed=n_buckets)
    tree = pkg.CellTree2d(vertices, faces, -1, n_buckets=n_buckets, cells_per_leaf=cells_per_leaf)
    ref = oracle.CellTree2d(vertices, faces, -1, n_buckets=n_buckets, cells_per_leaf=cells_per_leaf)
    assert_same_tree(tree, ref)


def test_build_structured_quads_bit_exact(pkg):
    vertices, faces = quad_mesh(512, 384)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    assert_same_tree(tree, ref)
    pts = np.random.default_rng(5).uniform(-0.05, 1.05, (500_000, 2))
    idx, w = tree.compute_barycentric_weights(pts)
    ridx, rw = ref.compute_barycentric_weights(pts)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(w, rw)


def test_points_and_weights(delaunay_pair):
    tree, ref, _, _ = delaunay_pair
    pts = np.random.default_rng(1).uniform(-0.02, 1.02, (2_000_000, 2))
    for tol in (None, 1e-7):
        idx, w = tree.compute_barycentric_weights(pts, tolerance=tol)
        ridx, rw = ref.compute_barycentric_weights(pts, tolerance=tol)
        assert np.array_equal(idx, ridx)
        np.testing.assert_allclose(w, rw, rtol=RTOL, atol=0)
        assert np.array_equal(w, rw)


def test_points_on_planes_and_entry_grid_boundaries(pkg):
    """Points exactly on split planes / cell corners / entry-grid cell boundaries, and one ulp to either side."""
    for nx, ny in ((64, 64), (96, 40)):
        vertices, faces = quad_mesh(nx, ny)
        tree = pkg.CellTree2d(vertices, faces, -1)
        ref = oracle.CellTree2d(vertices, faces, -1)
        xs = np.unique(np.concatenate([vertices[:, 0], 0.5 * (vertices[:-1, 0] + vertices[1:, 0])]))
        ys = np.unique(np.concatenate([vertices[:, 1], np.linspace(0, 1, 4 * ny + 1)]))
        xs = np.concatenate([xs, np.nextafter(xs, -np.inf), np.nextafter(xs, np.inf), [-0.25, 1.25, np.inf, -np.inf, np.nan]])
        ys = np.concatenate([ys, np.nextafter(ys, -np.inf), np.nextafter(ys, np.inf), [-0.25, 1.25, np.inf, -np.inf, np.nan]])
        xx, yy = np.meshgrid(xs, ys)
        points = np.column_stack((xx.ravel(), yy.ravel()))
        for tolerance in (None, 0.0, 1e-9):
            got = tree.locate_points(points, tolerance)
            want = ref.locate_points(points, tolerance)
            assert np.array_equal(got, want)
    # an irregular mesh: the same kinds of points around its vertex coordinates
    vertices, faces = delaunay_mesh(4_000, seed=21)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    rng = np.random.default_rng(8)
    pick = vertices[rng.integers(0, len(vertices), 20_000)]
    points = np.concatenate([pick, np.nextafter(pick, -np.inf), np.nextafter(pick, np.inf), np.column_stack((pick[:, 0], pick[::-1, 1]))])
    assert np.array_equal(tree.locate_points(points), ref.locate_points(points))
    i, w = tree.compute_barycentric_weights(points)
    ri, rw = ref.compute_barycentric_weights(points)
    assert np.array_equal(i, ri) and np.array_equal(w, rw)


def test_morton_ordering_is_invisible(pkg, delaunay_pair):
    from numba_celltree_b200 import _lib

    tree, _, _, _ = delaunay_pair
    pts = np.random.default_rng(2).uniform(0, 1, (1_500_000, 2))
    lib = _lib.load()
    results = []
    try:
        for bits in (0, 8, 16, 24, 32):
            lib.ct_set_sort_bits(bits)
            results.append(tree.compute_barycentric_weights(pts))
    finally:
        lib.ct_set_sort_bits(-1)
    for idx, w in results[1:]:
        assert np.array_equal(idx, results[0][0])
        assert np.array_equal(w, results[0][1])


def test_crowded_query_sets(pkg, delaunay_pair):
    """Points crowded into a few bins of the execution order (slabs that overflow), identical points, sorted points:
    the order is an execution detail, the answers are the oracle's."""
    tree, ref, _, _ = delaunay_pair
    rng = np.random.default_rng(12)
    n = 1_200_000
    crowded = np.column_stack((rng.normal(0.37, 0.002, n), rng.normal(0.61, 0.001, n)))  # nearly all in a handful of bins
    crowded[::1000] = rng.uniform(-0.1, 1.1, (len(crowded[::1000]), 2))
    crowded[5::997] = np.nan  # NaN and far-away points share the first bin of the order
    crowded[7::991] = [-1e300, 1e300]
    same = np.full((400_000, 2), 0.25)
    same[::3] = [0.75, 0.5]
    grid = np.stack(np.meshgrid(np.linspace(0, 1, 700), np.linspace(0, 1, 600), indexing="xy"), axis=-1).reshape(-1, 2)  # sorted rows
    for pts in (crowded, same, grid):
        i, w = tree.compute_barycentric_weights(pts)
        ri, rw = ref.compute_barycentric_weights(pts)
        assert np.array_equal(i, ri) and np.array_equal(w, rw)


def test_device_resident_inputs_match_host(delaunay_pair):
    torch = pytest.importorskip("torch")
    tree, _, _, faces = delaunay_pair
    pts = np.random.default_rng(3).uniform(0, 1, (300_000, 2))
    idx, w = tree.compute_barycentric_weights(pts)
    d_idx, d_w = tree.compute_barycentric_weights(torch.from_numpy(pts).cuda())
    assert d_idx.is_cuda and d_w.is_cuda
    assert np.array_equal(d_idx.cpu().numpy(), idx) and np.array_equal(d_w.cpu().numpy(), w)
    boxes = c3_boxes(len(faces), 50_000)
    i, j, a = tree.intersect_boxes(boxes)
    di, dj, da = tree.intersect_boxes(torch.from_numpy(boxes).cuda())
    assert np.array_equal(di.cpu().numpy(), i) and np.array_equal(dj.cpu().numpy(), j) and np.array_equal(da.cpu().numpy(), a)
    edges = c4_edges(len(faces), 20_000)
    i, j, xy = tree.intersect_edges(edges)
    di, dj, dxy = tree.intersect_edges(torch.from_numpy(edges).cuda())
    assert np.array_equal(di.cpu().numpy(), i) and np.array_equal(dj.cpu().numpy(), j)
    assert np.array_equal(dxy.cpu().numpy(), xy, equal_nan=True)


def test_device_resident_query_mesh_matches_host(delaunay_pair):
    """intersect_faces / locate_faces with the query mesh as CUDA tensors: same pairs and areas, results on the device,
    and locate_faces turns clockwise query faces counter-clockwise in place (celltree.py:212) on the device too."""
    torch = pytest.importorskip("torch")
    tree, ref, _, _ = delaunay_pair
    qv, qf = quad_mesh(150, 120)
    qf = qf.copy()
    qf[::3] = qf[::3, ::-1]  # every third face clockwise
    ri, rj, ra = ref.intersect_faces(qv, qf, -1)
    dqv, dqf = torch.from_numpy(qv).cuda(), torch.from_numpy(qf).cuda()
    di, dj, da = tree.intersect_faces(dqv, dqf, -1)
    assert di.is_cuda and dj.is_cuda and da.is_cuda
    assert np.array_equal(di.cpu().numpy(), ri) and np.array_equal(dj.cpu().numpy(), rj)
    np.testing.assert_allclose(da.cpu().numpy(), ra, rtol=RTOL, atol=0)
    assert np.array_equal(dqf.cpu().numpy(), qf), "intersect_faces must not modify the caller's faces"
    host_faces = qf.copy()
    hi, hj = tree.locate_faces(qv, host_faces)
    li, lj = tree.locate_faces(dqv, dqf)
    assert np.array_equal(li.cpu().numpy(), hi) and np.array_equal(lj.cpu().numpy(), hj)
    assert np.array_equal(dqf.cpu().numpy(), host_faces), "locate_faces makes the faces counter-clockwise in place"
    with pytest.raises(ValueError):
        tree.intersect_faces(dqv, qf, -1)  # mixed host / device mesh
    with pytest.raises(ValueError):
        tree.intersect_faces(dqv, dqf.to(torch.int32), -1)


def test_device_resident_network_queries_match_host(pkg):
    """EdgeCellTree2d with CUDA tensors: locate_points and intersect_edges give the host path's results, on the device."""
    torch = pytest.importorskip("torch")
    from numba_celltree_b200.synthetic import random_network

    vertices, edges = random_network(4000, seed=12)
    net = pkg.EdgeCellTree2d(vertices, edges)
    rng = np.random.default_rng(6)
    lo, hi = vertices.min(axis=0), vertices.max(axis=0)
    a = rng.uniform(lo, hi, (30_000, 2))
    segments = np.stack((a, a + rng.normal(0, 3.0, a.shape)), axis=1)
    i, j, xy = net.intersect_edges(segments)
    di, dj, dxy = net.intersect_edges(torch.from_numpy(segments).cuda())
    assert len(i) > 1000 and di.is_cuda and dxy.is_cuda and dxy.shape == (len(i), 2)
    assert np.array_equal(di.cpu().numpy(), i) and np.array_equal(dj.cpu().numpy(), j)
    assert np.array_equal(dxy.cpu().numpy(), xy, equal_nan=True)
    on_network = 0.5 * (vertices[edges[:, 0]] + vertices[edges[:, 1]])
    points = np.concatenate([on_network, rng.uniform(lo, hi, (20_000, 2))])
    found = net.locate_points(points)
    d_found = net.locate_points(torch.from_numpy(points).cuda())
    assert (found >= 0).sum() >= len(on_network) and np.array_equal(d_found.cpu().numpy(), found)
    with pytest.raises(ValueError):
        net.intersect_edges(torch.zeros((5, 4), dtype=torch.float64, device="cuda"))


def test_boxes(delaunay_pair):
    tree, ref, _, faces = delaunay_pair
    boxes = c3_boxes(len(faces), 300_000)
    i, j = tree.locate_boxes(boxes)
    ri, rj = ref.locate_boxes(boxes)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    i, j, a = tree.intersect_boxes(boxes)
    ri, rj, ra = ref.intersect_boxes(boxes)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    np.testing.assert_allclose(a, ra, rtol=RTOL, atol=0)
    assert np.array_equal(a, ra)


def test_edges(delaunay_pair):
    tree, ref, _, faces = delaunay_pair
    edges = c4_edges(len(faces), 100_000)
    i, j, xy = tree.intersect_edges(edges)
    ri, rj, rxy = ref.intersect_edges(edges)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    np.testing.assert_allclose(xy, rxy, rtol=RTOL, atol=0)
    assert np.array_equal(xy, rxy, equal_nan=True)


@pytest.mark.filterwarnings("ignore:overflow encountered", "ignore:invalid value encountered")  # t of a hit on an overflowing segment
def test_edges_with_extreme_coordinates(pkg, delaunay_pair):
    """Segments whose plane test leaves the finite case (infinite / overflowing / NaN coordinates), axis-parallel and
    degenerate segments, subnormal offsets: the division-free plane test must decide exactly as query.py:407-440 does.
    Both code paths of intersect_edges (log, and second traversal) and both tree kinds are covered."""
    from numba_celltree_b200 import _lib
    from numba_celltree_b200.synthetic import random_network

    tree, ref, vertices, faces = delaunay_pair
    rng = np.random.default_rng(17)
    n = 4000
    a = rng.uniform(-0.2, 1.2, (n, 2))
    b = rng.uniform(-0.2, 1.2, (n, 2))
    special = np.array([np.inf, -np.inf, np.nan, 1e308, -1e308, 1.7e308, 5e-324, -5e-324, 0.0, -0.0, 1e-300, 0.5, 2.0])
    which = rng.integers(0, 4, n)
    value = special[rng.integers(0, len(special), n)]
    a[which == 0, 0] = value[which == 0]
    a[which == 1, 1] = value[which == 1]
    b[which == 2, 0] = value[which == 2]
    b[which == 3, 1] = value[which == 3]
    both = rng.random(n) < 0.3  # a second special coordinate on the other end point
    b[both, 0] = special[rng.integers(0, len(special), int(both.sum()))]
    edges = np.stack((a, b), axis=1)
    axis = rng.uniform(0, 1, (2000, 2, 2))
    axis[:1000, 1, 0] = axis[:1000, 0, 0]  # vertical
    axis[1000:, 1, 1] = axis[1000:, 0, 1]  # horizontal
    on_vertices = np.stack((vertices[rng.integers(0, len(vertices), 1000)], vertices[rng.integers(0, len(vertices), 1000)]), axis=1)
    edges = np.concatenate([edges, axis, on_vertices, on_vertices[:, ::-1], np.repeat(on_vertices[:50, :1], 2, axis=1)])
    ri, rj, rxy = ref.intersect_edges(edges)
    assert len(ri) > 1000
    try:
        for per_query in (16, 0):
            _lib.check(_lib.load().ct_set_hit_log(per_query))
            i, j, xy = tree.intersect_edges(edges)
            assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(xy, rxy, equal_nan=True)
    finally:
        _lib.check(_lib.load().ct_set_hit_log(-1))
    net_vertices, net_edges = random_network(3000, seed=4)
    net = pkg.EdgeCellTree2d(net_vertices, net_edges)
    net_ref = oracle.EdgeCellTree2d(net_vertices, net_edges)
    lo, hi = net_vertices.min(), net_vertices.max()
    scaled = edges.copy()
    finite = np.isfinite(scaled) & (np.abs(scaled) < 1e3)
    scaled[finite] = lo + scaled[finite] * (hi - lo)
    i, j, xy = net.intersect_edges(scaled)
    ri, rj, rxy = net_ref.intersect_edges(scaled)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(xy, rxy, equal_nan=True)


def test_boxes_and_edges_with_a_short_or_absent_hit_log(pkg, delaunay_pair):
    """The hit log is an execution detail: overflowing it (or having none) falls back to the second traversal."""
    from numba_celltree_b200 import _lib

    tree, ref, _, faces = delaunay_pair
    edges = c4_edges(len(faces), 200_000)
    boxes = c3_boxes(len(faces), 200_000)
    ri, rj, rxy = ref.intersect_edges(edges)
    bi, bj, ba = ref.intersect_boxes(boxes)
    try:
        for per_query in (0, 1, 16):
            _lib.check(_lib.load().ct_set_hit_log(per_query))
            i, j, xy = tree.intersect_edges(edges)
            assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(xy, rxy, equal_nan=True)
            i, j, a = tree.intersect_boxes(boxes)
            assert np.array_equal(i, bi) and np.array_equal(j, bj) and np.array_equal(a, ba)
    finally:
        _lib.check(_lib.load().ct_set_hit_log(-1))


def test_faces_and_self_intersection(delaunay_pair):
    tree, ref, vertices, faces = delaunay_pair
    qv, qf = quad_mesh(150, 150)
    i, j, a = tree.intersect_faces(qv, qf, -1)
    ri, rj, ra = ref.intersect_faces(qv, qf, -1)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    np.testing.assert_allclose(a, ra, rtol=RTOL, atol=0)
    # the total overlap equals the area of the triangulated convex hull (size-independent property)
    tri = vertices[faces]
    u, w = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    hull_area = 0.5 * np.abs(u[:, 0] * w[:, 1] - u[:, 1] * w[:, 0]).sum()
    assert abs(a.sum() - hull_area) < 1e-9
    # the mesh against itself: sliver pairs at rounding-noise level must match too
    i, j, a = tree.intersect_faces(vertices, faces, -1)
    ri, rj, ra = ref.intersect_faces(vertices, faces, -1)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    assert np.array_equal(a, ra)


def test_intersect_faces_leaves_the_callers_faces_alone(delaunay_pair):
    """intersect_faces works on a copy (celltree.py:256): clockwise query faces with a foreign fill value stay as given."""
    tree, ref, _, _ = delaunay_pair
    qv, quads = quad_mesh(60, 60)
    # mixed mesh: every other quad becomes a triangle padded with -999, all faces clockwise
    qf = quads[:, ::-1].copy()
    qf[::2, 3] = -999
    given = qf.copy()
    i, j, a = tree.intersect_faces(qv, qf, -999)
    assert np.array_equal(qf, given)
    ri, rj, ra = ref.intersect_faces(qv, qf, -999)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra)
    # a non-contiguous / int32 faces array goes through cast_faces' copy and gives the same pairs
    i2, j2, a2 = tree.intersect_faces(qv, qf.astype(np.int32), -999)
    assert np.array_equal(i, i2) and np.array_equal(j, j2) and np.array_equal(a, a2)
    # locate_faces does rewrite its argument counter-clockwise (celltree.py:212)
    qf3 = np.where(qf == -999, -1, qf)
    rf3 = qf3.copy()
    pi, pj = tree.locate_faces(qv, qf3)
    rpi, rpj = ref.locate_faces(qv, rf3)
    assert np.array_equal(pi, rpi) and np.array_equal(pj, rpj)
    assert np.array_equal(qf3, rf3) and not np.array_equal(qf3, np.where(given == -999, -1, given))


def test_results_in_pinned_memory_equal_pageable_results(pkg, delaunay_pair):
    """Large results live in recycled page-locked blocks; content, dtype and writability are those of plain arrays."""
    from numba_celltree_b200 import _lib

    tree, _, _, faces = delaunay_pair
    boxes = c3_boxes(len(faces), 300_000)
    points = np.random.default_rng(3).uniform(0, 1, (400_000, 2))
    try:
        _lib.set_pinned_results(False)
        plain = tree.intersect_boxes(boxes) + tree.compute_barycentric_weights(points)
        _lib.set_pinned_results(True)
        for _ in range(2):  # the second round reuses the blocks released by the first
            pinned = tree.intersect_boxes(boxes) + tree.compute_barycentric_weights(points)
            for a, b in zip(plain, pinned):
                assert a.dtype == b.dtype and a.shape == b.shape and b.flags.writeable and b.flags.c_contiguous
                assert np.array_equal(a, b)
            assert pinned[0].nbytes >= _lib.PINNED_MIN_BYTES and not pinned[0].flags.owndata
            pinned[0][:] = 0  # writable like any result
            kept = pinned[2].copy()
            view = pinned[2]
            del pinned
            assert np.array_equal(view, kept)  # a surviving result keeps its block
    finally:
        _lib.set_pinned_results(True)
        _lib.load().ct_host_trim()


def test_save_and_load_round_trip(pkg, delaunay_pair, tmp_path):
    """A tree written to .npz and read back answers like the original (and refuses the other tree kind)."""
    from numba_celltree_b200.synthetic import random_network

    tree, ref, _, faces = delaunay_pair
    tree.save(tmp_path / "faces.npz")
    again = pkg.CellTree2d.load(tmp_path / "faces.npz")
    assert_same_tree(again, ref)
    points = np.random.default_rng(12).uniform(0, 1, (50_000, 2))
    assert np.array_equal(again.locate_points(points), ref.locate_points(points))
    edges = c4_edges(len(faces), 20_000)
    for a, b in zip(again.intersect_edges(edges), ref.intersect_edges(edges)):
        assert np.array_equal(a, b, equal_nan=True)
    with pytest.raises(ValueError):
        pkg.EdgeCellTree2d.load(tmp_path / "faces.npz")
    vertices, segments = random_network(5_000, seed=3)
    net = pkg.EdgeCellTree2d(vertices, segments)
    net.save(tmp_path / "network.npz")
    net2 = pkg.EdgeCellTree2d.load(tmp_path / "network.npz")
    assert np.array_equal(net2.nodes, net.nodes) and np.array_equal(net2.bb_indices, net.bb_indices)
    assert np.array_equal(net2.locate_points(vertices[:2000]), net.locate_points(vertices[:2000]))


def test_edge_network(pkg):
    from numba_celltree_b200.synthetic import random_network

    vertices, edges = random_network(50_000, seed=9)
    tree = pkg.EdgeCellTree2d(vertices, edges)
    ref = oracle.EdgeCellTree2d(vertices, edges)
    assert_same_tree(tree, ref)
    rng = np.random.default_rng(4)
    p, q = vertices[edges[:, 0]], vertices[edges[:, 1]]
    on = p + rng.uniform(0, 1, (len(edges), 1)) * (q - p)
    pts = np.concatenate([on, on + rng.normal(0, 1e-9, on.shape), rng.uniform(vertices.min(0), vertices.max(0), (100_000, 2))])
    for tol in (None, 1e-6):
        assert np.array_equal(tree.locate_points(pts, tolerance=tol), ref.locate_points(pts, tolerance=tol))
    a0 = rng.uniform(vertices.min(0), vertices.max(0), (50_000, 2))
    qe = np.stack((a0, a0 + rng.normal(0, 3.0, a0.shape)), axis=1)
    i, j, xy = tree.intersect_edges(qe)
    ri, rj, rxy = ref.intersect_edges(qe)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    assert np.array_equal(xy, rxy, equal_nan=True)


# ---- edge cases ------------------------------------------------------------------------------------------------
def test_empty_query_sets(delaunay_pair):
    tree, _, _, _ = delaunay_pair
    assert tree.locate_points(np.empty((0, 2))).shape == (0,)
    idx, w = tree.compute_barycentric_weights(np.empty((0, 2)))
    assert idx.shape == (0,) and w.shape == (0, 3)
    i, j = tree.locate_boxes(np.empty((0, 4)))
    assert i.shape == (0,) and j.shape == (0,) and i.dtype == np.intp
    i, j, a = tree.intersect_boxes(np.empty((0, 4)))
    assert i.shape == (0,) and a.shape == (0,)
    i, j, xy = tree.intersect_edges(np.empty((0, 2, 2)))
    assert i.shape == (0,) and xy.shape == (0, 2, 2)
    i, j, a = tree.intersect_faces(np.zeros((3, 2)), np.empty((0, 3), dtype=int), -1)
    assert i.shape == (0,) and a.shape == (0,)


def test_everything_outside_and_nan(delaunay_pair):
    tree, ref, _, _ = delaunay_pair
    pts = np.array([[5.0, 5.0], [-3.0, 0.5], [np.nan, 0.5], [0.5, np.nan], [np.inf, 0.5], [0.5, -np.inf]])
    assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))
    boxes = np.array([[2.0, 3.0, 2.0, 3.0], [np.nan, 1.0, 0.0, 1.0], [0.6, 0.4, 0.6, 0.4]])
    i, j = tree.locate_boxes(boxes)
    ri, rj = ref.locate_boxes(boxes)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    edges = np.array([[[2.0, 2.0], [3.0, 3.0]], [[0.5, 0.5], [0.5, 0.5]], [[np.nan, 0.5], [0.6, 0.5]]])
    i, j, xy = tree.intersect_edges(edges)
    ri, rj, rxy = ref.intersect_edges(edges)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(xy, rxy, equal_nan=True)


def test_single_cell_and_tiny_meshes(pkg):
    vertices = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]])
    for faces in ([[0, 1, 2]], [[0, 1, 2], [1, 3, 2]], [[0, 1, 3, 2]]):
        tree = pkg.CellTree2d(vertices, faces, -1)
        ref = oracle.CellTree2d(vertices, faces, -1)
        assert_same_tree(tree, ref)
        pts = np.random.default_rng(0).uniform(-0.5, 1.5, (1000, 2))
        assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))
    with pytest.raises(ValueError):
        pkg.CellTree2d(vertices, np.empty((0, 3), dtype=int), -1)


def geometric_strip(n):
    """n quads with x-extent halving each time: every split peels off the largest cells => a deep, thin tree."""
    edges_x = 1.0 / 2.0 ** np.arange(n + 1)
    vertices = np.concatenate([np.column_stack((edges_x, np.zeros(n + 1))), np.column_stack((edges_x, np.ones(n + 1)))])
    k = np.arange(n)
    faces = np.column_stack((k + 1, k, k + n + 1, k + n + 2))
    return vertices, faces


def nested_strip(n):
    """n quads [-x, 3x] x [0, 1] with x halving each time: all contain the origin and their centroids halve => a tree as
    deep as the strip whose nodes all overlap, so a query near the origin defers a sibling at every level."""
    x = 1.0 / 2.0 ** np.arange(n)
    k = np.arange(n)
    vertices = np.concatenate(
        [np.column_stack((-x, np.zeros(n))), np.column_stack((3 * x, np.zeros(n))), np.column_stack((3 * x, np.ones(n))), np.column_stack((-x, np.ones(n)))]
    )
    faces = np.column_stack((k, k + n, k + 2 * n, k + 3 * n))
    return vertices, faces


def test_trees_deeper_than_the_per_thread_stack_are_walked(pkg):
    """The reference's stack grows (utils.py:35-44): a tree of any depth is answered.  geometric_strip: deep but without
    overlap (the stack stays shallow); nested_strip: 299 levels with every node overlapping (stacks of ~300 entries)."""
    vertices, faces = geometric_strip(60)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    assert_same_tree(tree, ref)
    pts = np.column_stack((np.random.default_rng(0).uniform(0, 1, 5000) ** 8, np.full(5000, 0.5)))
    assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))
    vertices, faces = geometric_strip(400)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    assert_same_tree(tree, ref)
    assert tree.depth > 64
    assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))

    n = 300
    vertices, faces = nested_strip(n)
    tree = pkg.CellTree2d(vertices, faces, -1)
    ref = oracle.CellTree2d(vertices, faces, -1)
    assert_same_tree(tree, ref)
    assert tree.depth > 250
    rng = np.random.default_rng(1)
    m = 20_000
    pts = np.column_stack((rng.choice([-1.0, 1.0], m) * 2.0 ** -rng.uniform(0, n + 5, m), rng.uniform(-0.1, 1.1, m)))
    i, w = tree.compute_barycentric_weights(pts)
    ri, rw = ref.compute_barycentric_weights(pts)
    assert np.array_equal(i, ri) and np.array_equal(w, rw)
    assert len(np.unique(i)) > 200
    boxes = np.column_stack((-(2.0 ** -rng.uniform(0, n, 2000)), 2.0 ** -rng.uniform(0, n, 2000), rng.uniform(0, 0.4, 2000), rng.uniform(0.5, 1, 2000)))
    bi, bj = tree.locate_boxes(boxes)
    rbi, rbj = ref.locate_boxes(boxes)
    assert np.array_equal(bi, rbi) and np.array_equal(bj, rbj) and len(bi) > 100_000
    ai, aj, area = tree.intersect_boxes(boxes)
    rai, raj, rarea = ref.intersect_boxes(boxes)
    assert np.array_equal(ai, rai) and np.array_equal(aj, raj) and np.array_equal(area, rarea)
    a = np.column_stack((-(2.0 ** -rng.uniform(0, n, 1500)), rng.uniform(0, 1, 1500)))
    b = np.column_stack((2.0 ** -rng.uniform(0, n, 1500), rng.uniform(0, 1, 1500)))
    segments = np.stack((a, b), axis=1)
    ei, ej, xy = tree.intersect_edges(segments)
    rei, rej, rxy = ref.intersect_edges(segments)
    assert np.array_equal(ei, rei) and np.array_equal(ej, rej) and np.array_equal(xy, rxy, equal_nan=True)
    # device-resident, Morton / binned order (a batch large enough to be ordered)
    big = np.column_stack((rng.choice([-1.0, 1.0], 400_000) * 2.0 ** -rng.uniform(0, n + 5, 400_000), rng.uniform(0, 1, 400_000)))
    assert np.array_equal(tree.locate_points(big), ref.locate_points(big))


def test_unbucketable_centroid_raises_like_the_reference(pkg):
    # a zero-width cell at the upper end of the range falls in no bucket: the reference indexes past its
    # bucket list (IndexError, creation.py:130-131); the oracle and the CUDA build raise IndexError too
    vertices = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0], [2.0, 0.0], [2.0, 1.0], [2.0, 0.5]])
    faces = np.array([[0, 1, 2, 3], [1, 4, 5, 2], [4, 5, 6, -1]])
    with pytest.raises(IndexError):
        oracle.CellTree2d(vertices, faces, -1, cells_per_leaf=1)
    with pytest.raises(IndexError):
        pkg.CellTree2d(vertices, faces, -1, cells_per_leaf=1)


def mixed_mesh(n_points, seed):
    """Triangles, quads (pairs of Delaunay triangles merged) and padded rows in one face array of width 4."""
    vertices, tri = delaunay_mesh(n_points, seed)
    faces = np.full((len(tri), 4), -1, dtype=tri.dtype)
    faces[:, :3] = tri
    return vertices, faces


def hexagon_mesh(nx, ny):
    """Rows of regular hexagons (six vertices a face, the generic polygon path), every other row shifted."""
    angles = np.pi / 3 * np.arange(6)
    ring = np.column_stack((np.cos(angles), np.sin(angles)))
    centres = np.array([(1.5 * i, np.sqrt(3) * (j + 0.5 * (i % 2))) for j in range(ny) for i in range(nx)])
    corners = (centres[:, None, :] + ring[None, :, :]).reshape(-1, 2)
    unique, inverse = np.unique(np.round(corners, 9), axis=0, return_inverse=True)
    return unique, inverse.reshape(-1, 6).astype(np.int64)


@pytest.mark.filterwarnings("ignore:overflow encountered", "ignore:invalid value encountered")
def test_the_bounding_filter_never_changes_an_answer(pkg):
    """The traversal skips a cell's point-in-polygon test when the point is surely outside the cell's bounds
    (geometry.cuh: point_surely_outside); the reference tests every cell of a leaf.  Tolerances from zero to infinity,
    negative and NaN, points on and a few ulps around the cells' bounding lines and on the extensions of their edges,
    coordinates shifted to 1e7 (tolerance below one ulp), scaled to 1e-160 / 1e-300 (products underflow) and to 1e150 /
    1e300 (products overflow)."""
    rng = np.random.default_rng(77)
    meshes = [delaunay_mesh(3_000, seed=5), mixed_mesh(2_000, seed=6), quad_mesh(40, 30), hexagon_mesh(12, 10)]
    for vertices, faces in meshes:
        lo, hi = vertices.min(0), vertices.max(0)
        span = hi - lo
        n_vert = (faces >= 0).sum(1)
        first = vertices[faces[:, 0]]
        second = vertices[faces[:, 1]]
        t = rng.uniform(-1.5, 2.5, (len(faces), 1))
        on_lines = first + t * (second - first)  # on the (extended) line through an edge of every cell
        pick = vertices[rng.integers(0, len(vertices), 3_000)]
        cross = np.column_stack((pick[:, 0], pick[::-1, 1]))  # on the bounding lines of two different cells
        base = np.concatenate([rng.uniform(lo - 0.1 * span, hi + 0.1 * span, (10_000, 2)), on_lines, pick, cross])
        points = np.concatenate([base, np.nextafter(base, -np.inf), np.nextafter(base, np.inf), base + 1e-9, base - 3e-7])
        for shift, scale in ((0.0, 1.0), (1e7, 1.0), (0.0, 1e-160), (0.0, 1e-300), (0.0, 1e150), (0.0, 1e300), (-3e5, 37.0)):
            v = (vertices + shift) * scale
            q = (points + shift) * scale
            every = (None, 0.0, 1e-9 * scale, 0.02 * span[0] * scale, 3.0 * span[0] * scale, -1e-3 * scale, np.inf, np.nan)
            for cells_per_leaf, tolerances in ((2, every), (5, every[::3])):
                tree = pkg.CellTree2d(v, faces, -1, cells_per_leaf=cells_per_leaf)
                ref = oracle.CellTree2d(v, faces, -1, cells_per_leaf=cells_per_leaf)
                for tolerance in tolerances:
                    got = tree.locate_points(q, tolerance)
                    want = ref.locate_points(q, tolerance)
                    assert np.array_equal(got, want), (shift, scale, tolerance, int(n_vert.max()))


ORDER_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, {root!r})
import oracle
from numba_celltree_b200 import CellTree2d, _lib
from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh
_lib.load().ct_set_sort_bits(16)  # the binned path also for these small batches
rng = np.random.default_rng(3)
for vertices, faces in (delaunay_mesh(20_000, seed=9), quad_mesh(300, 200)):
    tree, ref = CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1)
    uniform = rng.uniform(-0.05, 1.05, (30_000, 2))
    crowded = np.concatenate([uniform[:7_000], rng.normal(0.4, 0.0005, (9_000, 2)), np.full((3_000, 2), 0.7)])
    sparse = np.concatenate([rng.uniform(0, 1, (300, 2)), [[np.nan, 0.5], [1e300, -1e300]]])
    for points in (uniform, crowded, sparse, uniform[:2048], uniform[:2049], uniform[:1]):
        got, want = tree.compute_barycentric_weights(points), ref.compute_barycentric_weights(points)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), len(points)
        assert np.array_equal(tree.locate_points(points), want[0])
    # leaves of more than two cells: the kernels that pick the candidate cell by its bounds first
    tree, ref = CellTree2d(vertices, faces, -1, cells_per_leaf=6), oracle.CellTree2d(vertices, faces, -1, cells_per_leaf=6)
    for points in (uniform, crowded):
        for tolerance in (None, 0.0, 0.01):
            assert np.array_equal(tree.locate_points(points, tolerance), ref.locate_points(points, tolerance))
print("same answers")
"""


@pytest.mark.parametrize("order", ["slabs", "bins", "sort", "morton"])
def test_every_execution_order_gives_the_oracles_answers(order):
    """CELLTREE_ORDER is read once per process, so every order runs in a process of its own: full, partial and
    single-point tiles, crowded and sparse batches through the binned kernels."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CELLTREE_ORDER=order)
    done = subprocess.run([sys.executable, "-c", ORDER_SCRIPT.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
    assert done.returncode == 0 and "same answers" in done.stdout, done.stdout[-2000:] + done.stderr[-4000:]
