"""CPU: the oracle's restatement of the exported geometry helpers, and the host-side demo helpers, against the golden
vectors the reference itself produced (tests/golden/make_golden_algorithms.py)."""

import pathlib

import numpy as np
import pytest

import oracle

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "algorithms.npz"


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name, fn", [("lb", "liang_barsky_line_box_clip"), ("cs", "cohen_sutherland_line_box_clip_batch")])
def test_line_box_clip(g, name, fn):
    hit, c, d = getattr(oracle, fn)(g["seg_a"], g["seg_b"], g["seg_boxes"])
    assert same(hit, g[f"{name}_hit"]) and same(c, g[f"{name}_c"]) and same(d, g[f"{name}_d"])


@pytest.mark.parametrize("poly", ["square", "triangle", "hexagon"])
@pytest.mark.parametrize("tol", ["tol", "zero"])
def test_line_polygon_clip(g, poly, tol):
    k = f"cb_{poly}_{tol}"
    hit, c, d = oracle.cyrus_beck_line_polygon_clip_batch(g[f"{k}_a"], g[f"{k}_b"], g[f"{k}_poly"], float(g[f"{k}_tolerance"]))
    assert same(hit, g[f"{k}_hit"]) and same(c, g[f"{k}_c"]) and same(d, g[f"{k}_d"])


@pytest.mark.parametrize("poly", ["square", "triangle", "hexagon", "unit_square", "square_reversed", "concave", "repeated_vertex"])
def test_points_in_polygon(g, poly):
    assert same(oracle.points_in_polygon(g[f"pip_{poly}_points"], g[f"pip_{poly}_poly"]), g[f"pip_{poly}_inside"])


@pytest.mark.parametrize("tol", ["1e-09", "0.001", "0"])
def test_points_in_triangles(g, tol):
    k = f"pit_{tol}"
    got = oracle.points_in_triangles(g[f"{k}_points"], g[f"{k}_face_indices"], g["pit_faces"], g["pit_vertices"], float(g[f"{k}_tolerance"]))
    assert same(got, g[f"{k}_inside"])


def test_reference_known_answers():
    """tests/test_algorithms/test_line_box_clip.py:47-60 and tests/test_geometry_utils.py:130-146."""
    hit, c, d = oracle.liang_barsky_line_box_clip([[-1.0, 0.0]], [[2.0, 3.0]], [0.0, 2.0, 0.0, 2.0])
    assert hit[0] and np.allclose(c[0], [0.0, 1.0]) and np.allclose(d[0], [1.0, 2.0])
    hit, c, d = oracle.liang_barsky_line_box_clip([[0.0, -0.1]], [[0.0, -0.1]], [0.0, 2.0, 0.0, 2.0])
    assert not hit[0] and np.isnan(c).all() and np.isnan(d).all()
    poly = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    pts = [[0.5, 0.25], [1.5, 0.25], [0.0, 0.0], [0.5, 0.5], [1.0, 1.0]]
    assert oracle.points_in_polygon(pts, poly).tolist() == [True, False, True, True, False]
    assert oracle.points_in_polygon(pts, poly[::-1]).tolist() == [True, False, True, True, False]


def test_demo_helpers(g):
    from numba_celltree_b200 import demo

    assert same(demo.close_polygons(g["demo_faces"], -1), g["demo_closed"])
    assert same(demo.edges(g["demo_faces"], -1), g["demo_edges"])
    v, e = demo.example_1d_network()
    assert same(v, g["demo_network_vertices"]) and same(e, g["demo_network_edges"]) and e.dtype == g["demo_network_edges"].dtype
    xy, tri = demo.generate_disk(5, 5)
    assert xy.shape == (76, 2) and tri.shape[1] == 3
    with pytest.raises(ValueError):
        demo.generate_disk(2, 3)
