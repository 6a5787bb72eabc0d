"""
Shared parity checkers: run one implementation (the CPU oracle or the CUDA package) over a golden
fixture produced by the reference (tests/golden/make_golden.py) and compare.

Bar: indices, pair lists (including order), nodes and bb_indices bit-exact; areas, clip coordinates and
weights within RTOL = 1e-12 relative (the tolerance BASELINE.json's north_star states).  In practice
every float below also matches bit-for-bit; ``exact_floats=True`` asserts that too.
"""

import pathlib

import numpy as np

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden"
RTOL = 1e-12
BUILD_PARAMS = [(4, 2), (2, 1), (2, 2), (8, 3), (3, 1), (16, 4)]
FACE_CASES = ["disk_5_5", "triangles_538", "voronoi_74", "quads_48_40_mixed", "delaunay_3000", "duplicates", "lattice_faces"]
EDGE_CASES = ["network_demo", "network_800", "network_grid", "lattice_network"]


def load(name):
    return np.load(GOLDEN / f"{name}.npz")


def assert_close(actual, expected, what, exact_floats=True):
    actual = np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} != {expected.shape}"
    if exact_floats:
        same = (actual == expected) | (np.isnan(actual) & np.isnan(expected))
        assert same.all(), f"{what}: {np.count_nonzero(~same)} of {same.size} values differ bit-wise"
    np.testing.assert_allclose(actual, expected, rtol=RTOL, atol=0.0, equal_nan=True, err_msg=what)


def assert_same_nodes(actual, expected, what):
    assert actual.dtype == expected.dtype, f"{what}: dtype {actual.dtype} != {expected.dtype}"
    assert actual.shape == expected.shape, f"{what}: {actual.shape} nodes != {expected.shape}"
    for field in ("child", "Lmax", "Rmin", "ptr", "size", "dim"):
        assert np.array_equal(actual[field], expected[field]), f"{what}: nodes[{field!r}] differs"


def check_face_tree_build(CellTree2d, name):
    g = load(name)
    fill = int(g["fill_value"])
    for nb_, cpl in BUILD_PARAMS:
        t = CellTree2d(g["vertices"], g["faces"], fill, n_buckets=nb_, cells_per_leaf=cpl)
        p = f"b{nb_}_c{cpl}_"
        assert_same_nodes(t.nodes, g[p + "nodes"], f"{name} {p}")
        assert np.array_equal(t.bb_indices, g[p + "bb_indices"]), f"{name} {p} bb_indices"
        assert t.bb_indices.dtype == np.intp
        assert np.array_equal(t.bb_coords, g[p + "bb_coords"]), f"{name} {p} bb_coords"
        assert np.array_equal(t.bbox, g[p + "bbox"]), f"{name} {p} bbox"
        assert np.array_equal(t.faces, g["faces_ccw"]), f"{name} {p} faces after counter_clockwise"
    t = CellTree2d(g["vertices"], g["faces"], fill)
    assert np.array_equal(t.bb_distances, g["bb_distances"])


def check_face_tree_points(CellTree2d, name):
    g = load(name)
    t = CellTree2d(g["vertices"], g["faces"], int(g["fill_value"]))
    pts = g["points"]
    r = t.locate_points(pts)
    assert r.dtype == np.intp
    assert np.array_equal(r, g["locate_points"]), f"{name} locate_points"
    for k in (0, 1):
        r = t.locate_points(pts, tolerance=float(g[f"tol{k}"]))
        assert np.array_equal(r, g[f"locate_points_tol{k}"]), f"{name} locate_points tol{k}"
    fi, w = t.compute_barycentric_weights(pts)
    assert np.array_equal(fi, g["locate_points"])
    assert_close(w, g["weights"], f"{name} barycentric weights")
    if "weights_tol1" in g:
        fi, w = t.compute_barycentric_weights(pts, tolerance=float(g["tol1"]))
        assert np.array_equal(fi, g["locate_points_tol1"])
        assert_close(w, g["weights_tol1"], f"{name} barycentric weights tol1")


def check_face_tree_boxes(CellTree2d, name):
    g = load(name)
    t = CellTree2d(g["vertices"], g["faces"], int(g["fill_value"]))
    i, j = t.locate_boxes(g["boxes"])
    assert i.dtype == np.intp and j.dtype == np.intp
    assert np.array_equal(i, g["locate_boxes_i"]) and np.array_equal(j, g["locate_boxes_j"]), f"{name} locate_boxes"
    i, j, a = t.intersect_boxes(g["boxes"])
    assert np.array_equal(i, g["intersect_boxes_i"]) and np.array_equal(j, g["intersect_boxes_j"]), f"{name} intersect_boxes"
    assert_close(a, g["intersect_boxes_area"], f"{name} intersect_boxes area")


def check_face_tree_edges(CellTree2d, name):
    g = load(name)
    t = CellTree2d(g["vertices"], g["faces"], int(g["fill_value"]))
    i, j, xy = t.intersect_edges(g["edges"])
    assert np.array_equal(i, g["intersect_edges_i"]), f"{name} intersect_edges i"
    assert np.array_equal(j, g["intersect_edges_j"]), f"{name} intersect_edges j"
    assert_close(xy, g["intersect_edges_xy"], f"{name} intersect_edges xy")


def check_face_tree_faces(CellTree2d, name):
    g = load(name)
    t = CellTree2d(g["vertices"], g["faces"], int(g["fill_value"]))
    ov, of, ofill = g["other_vertices"], g["other_faces"], int(g["other_fill"])
    i, j, a = t.intersect_faces(ov, of, ofill)
    assert np.array_equal(i, g["intersect_faces_i"]) and np.array_equal(j, g["intersect_faces_j"]), f"{name} intersect_faces"
    assert_close(a, g["intersect_faces_area"], f"{name} intersect_faces area")
    of_c = np.where(of == ofill, -1, of).astype(np.intp)
    i, j = t.locate_faces(ov.copy(), of_c)
    assert np.array_equal(i, g["locate_faces_i"]) and np.array_equal(j, g["locate_faces_j"]), f"{name} locate_faces"
    i, j, a = t.intersect_faces(g["vertices"], g["faces"], int(g["fill_value"]))
    assert np.array_equal(i, g["self_faces_i"]) and np.array_equal(j, g["self_faces_j"]), f"{name} self intersect_faces"
    assert_close(a, g["self_faces_area"], f"{name} self intersect_faces area")


def check_edge_tree(EdgeCellTree2d, name):
    g = load(name)
    for nb_, cpl in BUILD_PARAMS:
        t = EdgeCellTree2d(g["vertices"], g["edges"], n_buckets=nb_, cells_per_leaf=cpl)
        p = f"b{nb_}_c{cpl}_"
        assert_same_nodes(t.nodes, g[p + "nodes"], f"{name} {p}")
        assert np.array_equal(t.bb_indices, g[p + "bb_indices"]), f"{name} {p} bb_indices"
        assert np.array_equal(t.bb_coords, g[p + "bb_coords"]), f"{name} {p} bb_coords"
        assert np.array_equal(t.bbox, g[p + "bbox"]), f"{name} {p} bbox"
    t = EdgeCellTree2d(g["vertices"], g["edges"])
    assert np.array_equal(t.bb_distances, g["bb_distances"])
    pts = g["points"]
    assert np.array_equal(t.locate_points(pts), g["locate_points"]), f"{name} locate_points"
    for k in (0, 1):
        r = t.locate_points(pts, tolerance=float(g[f"tol{k}"]))
        assert np.array_equal(r, g[f"locate_points_tol{k}"]), f"{name} locate_points tol{k}"
    i, j, xy = t.intersect_edges(g["query_edges"])
    assert np.array_equal(i, g["intersect_edges_i"]), f"{name} intersect_edges i"
    assert np.array_equal(j, g["intersect_edges_j"]), f"{name} intersect_edges j"
    assert_close(xy, g["intersect_edges_xy"], f"{name} intersect_edges xy")


def check_extreme_segments(CellTree2d, EdgeCellTree2d):
    """`extreme_segments.npz` (tests/golden/make_golden_extreme.py): infinite / NaN / overflowing / subnormal / axis-parallel
    query segments through intersect_edges of both tree kinds, against the reference's own output."""
    g = load("extreme_segments")
    i, j, xy = CellTree2d(g["face_vertices"], g["face_faces"], -1).intersect_edges(g["face_segments"])
    assert np.array_equal(i, g["face_i"]) and np.array_equal(j, g["face_j"]), "extreme segments: face pairs"
    assert_close(xy, g["face_xy"], "extreme segments: face clip coordinates")
    i, j, xy = EdgeCellTree2d(g["net_vertices"], g["net_edges"]).intersect_edges(g["net_segments"])
    assert np.array_equal(i, g["net_i"]) and np.array_equal(j, g["net_j"]), "extreme segments: network pairs"
    assert_close(xy, g["net_xy"], "extreme segments: network intersection points")
