"""GPU: behaviour at the boundary that the reference's users rely on or that CUDA tensors bring with them -- edits of
``tree.nodes`` reach the queries (the reference reads that array on every call), device-resident calls run on the
caller's stream, unaligned views work, tensors of another GPU are refused."""

import numpy as np
import pytest

import oracle
from numba_celltree_b200.synthetic import c3_boxes, delaunay_mesh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    import numba_celltree_b200 as pkg

    vertices, faces = delaunay_mesh(20_000, seed=99)
    return pkg.CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1), faces


def test_edits_of_the_node_array_reach_the_queries(pair):
    """tests/test_celltree.py:606-618 edits nodes[71]["Lmax"]; query.py:73 reads tree.nodes on every call."""
    import numba_celltree_b200 as pkg

    _, _, faces = pair
    vertices, faces = delaunay_mesh(20_000, seed=99)
    tree, ref = pkg.CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1)
    pts = np.random.default_rng(1).uniform(0, 1, (200_000, 2))
    boxes = c3_boxes(len(faces), 20_000)
    before = tree.locate_points(pts)
    assert np.array_equal(before, ref.locate_points(pts))
    # pull the left plane of a few upper inner nodes far to the left: their left subtrees are no longer reached
    inner = np.flatnonzero(ref.nodes["child"] != -1)[[1, 2, 5, 40]]
    for nodes in (tree.nodes, ref.nodes):
        nodes["Lmax"][inner] = -1.0
    after = tree.locate_points(pts)
    assert np.array_equal(after, ref.locate_points(pts))
    assert (after != before).sum() > 1000 and (after == -1).sum() > (before == -1).sum()
    i, j = tree.locate_boxes(boxes)
    ri, rj = ref.locate_boxes(boxes)
    assert np.array_equal(i, ri) and np.array_equal(j, rj)
    assert not tree.validate_node_bounds().all()
    # a second edit, through a record view this time (the reference's own test writes nodes[k]["Lmax"])
    tree.nodes[int(inner[0])]["Lmax"] = 2.0
    ref.nodes[int(inner[0])]["Lmax"] = 2.0
    assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))
    # links that do not form a tree are refused and the tree keeps working with what it had
    good = int(tree.nodes["child"][0])
    tree.nodes["child"][0] = len(tree.nodes) + 5
    with pytest.raises(ValueError):
        tree.locate_points(pts)
    tree.nodes["child"][0] = good
    assert np.array_equal(tree.locate_points(pts), ref.locate_points(pts))


def test_unaligned_views_and_side_streams(pair):
    torch = pytest.importorskip("torch")
    tree, ref, faces = pair
    n = 400_000
    pts = np.random.default_rng(5).uniform(0, 1, (n, 2))
    want_i, want_w = ref.compute_barycentric_weights(pts)
    flat = torch.empty(2 * n + 1, dtype=torch.float64, device="cuda")
    flat[1:] = torch.from_numpy(pts).cuda().reshape(-1)
    view = flat[1:].view(n, 2)  # starts 8 bytes into the buffer
    assert view.data_ptr() % 16 == 8
    i, w = tree.compute_barycentric_weights(view)
    assert np.array_equal(i.cpu().numpy(), want_i) and np.array_equal(w.cpu().numpy(), want_w)
    boxes = c3_boxes(len(faces), 30_001)
    bflat = torch.empty(4 * len(boxes) + 1, dtype=torch.float64, device="cuda")
    bflat[1:] = torch.from_numpy(boxes).cuda().reshape(-1)
    bi, bj = tree.locate_boxes(bflat[1:].view(-1, 4))
    ri, rj = ref.locate_boxes(boxes)
    assert np.array_equal(bi.cpu().numpy(), ri) and np.array_equal(bj.cpu().numpy(), rj)
    # produced, queried and consumed on a side stream: no synchronisation with the default stream in between
    side = torch.cuda.Stream()
    host = torch.from_numpy(pts).pin_memory()
    with torch.cuda.stream(side):
        for _ in range(3):
            dev = host.to("cuda", non_blocking=True) * 1.0  # a kernel on `side` produces the input
            found = tree.locate_points(dev)
            total = (found >= 0).sum()  # consumed on `side`
            del dev
        got = found.cpu()
    side.synchronize()
    assert np.array_equal(got.numpy(), want_i) and int(total.item()) == int((want_i >= 0).sum())


def test_tensors_of_another_gpu_are_refused_and_the_current_device_is_kept():
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import numba_celltree_b200 as pkg

    vertices, faces = delaunay_mesh(5_000, seed=3)
    tree = pkg.CellTree2d(vertices, faces, -1)  # on device 0
    pts = np.random.default_rng(0).uniform(0, 1, (10_000, 2))
    with pytest.raises(ValueError, match="lives on"):
        tree.locate_points(torch.from_numpy(pts).to("cuda:1"))
    torch.cuda.set_device(1)
    try:
        want = oracle.CellTree2d(vertices, faces, -1).locate_points(pts)
        assert np.array_equal(tree.locate_points(pts), want)  # host buffers, tree on device 0, current device 1
        assert torch.cuda.current_device() == 1
        assert np.array_equal(tree.locate_points(torch.from_numpy(pts).to("cuda:0")).cpu().numpy(), want)
        assert torch.cuda.current_device() == 1
    finally:
        torch.cuda.set_device(0)


def test_concave_cells_overflowing_the_small_clip_buffers():
    """The clip's working polygons hold MAXA + MAXB vertices, enough for convex cells; a concave or self-intersecting cell
    can make the clipped polygon grow beyond that, and the pair is then repeated with the reference's capacity (64)."""
    import numba_celltree_b200 as pkg

    rng = np.random.default_rng(8)
    n = 30
    xs, ys = np.meshgrid(np.arange(n, dtype=float), np.arange(n, dtype=float), indexing="xy")
    base = np.column_stack((xs.ravel(), ys.ravel()))
    # darts (concave quads) and bow ties (self-intersecting quads), counter-clockwise where that means anything
    dart = np.array([[0.0, 0.0], [0.9, 0.45], [0.0, 0.9], [0.35, 0.45]])
    bow = np.array([[0.0, 0.0], [0.9, 0.9], [0.9, 0.0], [0.0, 0.9]])
    shapes = np.where((rng.random(len(base)) < 0.5)[:, None, None], dart, bow)
    vertices = (base[:, None, :] + shapes).reshape(-1, 2)
    faces = np.arange(len(vertices)).reshape(-1, 4)
    tree, ref = pkg.CellTree2d(vertices, faces, -1), oracle.CellTree2d(vertices, faces, -1)
    c = rng.uniform(0, n, (20_000, 2))
    wh = rng.uniform(0.05, 2.5, (20_000, 2))
    boxes = np.column_stack((c[:, 0] - wh[:, 0] / 2, c[:, 0] + wh[:, 0] / 2, c[:, 1] - wh[:, 1] / 2, c[:, 1] + wh[:, 1] / 2))
    i, j, a = tree.intersect_boxes(boxes)
    ri, rj, ra = ref.intersect_boxes(boxes)
    assert np.array_equal(i, ri) and np.array_equal(j, rj) and np.array_equal(a, ra) and len(i) > 10_000
    # the same cells as a query mesh over themselves, shifted: concave subject AND concave clipper
    shifted = vertices + [0.3, 0.2]
    fi, fj, fa = tree.intersect_faces(shifted, faces, -1)
    rfi, rfj, rfa = ref.intersect_faces(shifted, faces, -1)
    assert np.array_equal(fi, rfi) and np.array_equal(fj, rfj) and np.array_equal(fa, rfa) and len(fi) > 300


def test_division_by_zero_in_the_weights():
    """Wachspress weights with tolerance=0.0 for a point exactly on an edge divide by zero.  The reference raises
    ZeroDivisionError inside its prange loop, with undefined consequences (measured: the exception for a batch of one, a
    normal return with part of the rows left zero for larger batches); the library always raises.  A triangle of zero area
    gives NaN weights in both."""
    import numba_celltree_b200 as pkg
    from numba_celltree_b200.synthetic import quad_mesh

    v, f = quad_mesh(4, 4)
    tree = pkg.CellTree2d(v, f, -1)
    pts = np.random.default_rng(0).uniform(0, 1, (1000, 2))
    i, w = tree.compute_barycentric_weights(pts, tolerance=0.0)  # no point on an edge: fine
    assert np.allclose(w.sum(axis=1), 1.0)
    pts[500] = [0.25, 0.1]
    with pytest.raises(ZeroDivisionError):
        tree.compute_barycentric_weights(pts, tolerance=0.0)
    i, w = tree.compute_barycentric_weights(pts)  # default tolerance: the on-edge interpolation takes over
    assert i[500] == 0 and np.array_equal(w[500], [0.0, 0.6, 0.4, 0.0])
    assert np.array_equal(tree.locate_points(pts, tolerance=0.0), oracle.CellTree2d(v, f, -1).locate_points(pts, 0.0))
    v3 = np.array([[0.0, 0.0], [1.0, 0.0], [2.0, 0.0], [0.0, 1.0]])
    f3 = np.array([[0, 1, 2], [0, 1, 3]])
    i3, w3 = pkg.CellTree2d(v3, f3, -1).compute_barycentric_weights(np.array([[0.5, 0.0], [0.2, 0.2]]))
    assert i3.tolist() == [0, 1] and np.isnan(w3[0]).all() and np.allclose(w3[1], [0.6, 0.2, 0.2], rtol=1e-15)
