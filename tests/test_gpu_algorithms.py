"""GPU: the batched C-ABI entries for the exported geometry helpers (algorithms.cu) against the reference's golden
vectors and the oracle, bit for bit; plus the scalar call forms the reference's users know."""

import pathlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = pathlib.Path(__file__).resolve().parent / "golden" / "algorithms.npz"


@pytest.fixture(scope="module")
def g():
    with np.load(GOLDEN) as z:
        return {k: z[k] for k in z.files}


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name, fn", [("lb", "liang_barsky_line_box_clip"), ("cs", "cohen_sutherland_line_box_clip")])
def test_line_box_clip(g, name, fn):
    from numba_celltree_b200 import algorithms

    hit, c, d = getattr(algorithms, fn)(g["seg_a"], g["seg_b"], g["seg_boxes"])
    assert same(hit, g[f"{name}_hit"]) and same(c, g[f"{name}_c"]) and same(d, g[f"{name}_d"])
    # one box for all segments
    box = (0.0, 2.0, 0.0, 2.0)
    rows = np.flatnonzero((g["seg_boxes"] == box).all(axis=1))
    hit1, c1, d1 = getattr(algorithms, fn)(g["seg_a"][rows], g["seg_b"][rows], algorithms.Box(*box))
    assert same(hit1, g[f"{name}_hit"][rows]) and same(c1, g[f"{name}_c"][rows]) and same(d1, g[f"{name}_d"][rows])


@pytest.mark.parametrize("poly", ["square", "triangle", "hexagon"])
@pytest.mark.parametrize("tol", ["tol", "zero"])
def test_line_polygon_clip(g, poly, tol):
    from numba_celltree_b200 import algorithms

    k = f"cb_{poly}_{tol}"
    hit, c, d = algorithms.cyrus_beck_line_polygon_clip(g[f"{k}_a"], g[f"{k}_b"], g[f"{k}_poly"], float(g[f"{k}_tolerance"]))
    assert same(hit, g[f"{k}_hit"]) and same(c, g[f"{k}_c"]) and same(d, g[f"{k}_d"])


@pytest.mark.parametrize("poly", ["square", "triangle", "hexagon", "unit_square", "square_reversed", "concave", "repeated_vertex"])
def test_points_in_polygon(g, poly):
    from numba_celltree_b200 import algorithms

    assert same(algorithms.points_in_polygon(g[f"pip_{poly}_points"], g[f"pip_{poly}_poly"]), g[f"pip_{poly}_inside"])


@pytest.mark.parametrize("tol", ["1e-09", "0.001", "0"])
def test_points_in_triangles(g, tol):
    from numba_celltree_b200 import algorithms

    k = f"pit_{tol}"
    got = algorithms.points_in_triangles(g[f"{k}_points"], g[f"{k}_face_indices"], g["pit_faces"], g["pit_vertices"], float(g[f"{k}_tolerance"]))
    assert same(got, g[f"{k}_inside"])
    with pytest.raises(ValueError):
        algorithms.points_in_triangles(g[f"{k}_points"][:4], [0, 1, 10**6, 2], g["pit_faces"], g["pit_vertices"], 0.0)


def test_scalar_forms_of_the_reference():
    """tests/test_algorithms/test_line_box_clip.py:47-60, tests/test_geometry_utils.py:130-146, :422-440 call shapes."""
    from numba_celltree_b200.algorithms import (
        Box, Point, Triangle, cohen_sutherland_line_box_clip, cyrus_beck_line_polygon_clip, liang_barsky_line_box_clip,
        point_in_polygon, point_in_triangle,
    )  # fmt: skip

    box = Box(0.0, 2.0, 0.0, 2.0)
    poly = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0]])
    for clip in (cohen_sutherland_line_box_clip, liang_barsky_line_box_clip, lambda a, b, _: cyrus_beck_line_polygon_clip(a, b, poly, 1e-9)):
        ok, c, d = clip(Point(-1.0, 0.0), Point(2.0, 3.0), box)
        assert ok is True and np.allclose(c, [0.0, 1.0]) and np.allclose(d, [1.0, 2.0]) and isinstance(c, Point)
        ok, c, d = clip(Point(0.0, -0.1), Point(0.0, -0.1), box)
        assert ok is False and np.isnan(c).all() and np.isnan(d).all()
    unit = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    assert point_in_polygon(Point(0.5, 0.25), unit) and not point_in_polygon(Point(1.5, 0.25), unit)
    assert point_in_polygon(Point(0.0, 0.0), unit) and not point_in_polygon(Point(1.0, 1.0), unit[::-1])
    t = Triangle(Point(0.1, 0.1), Point(0.7, 0.5), Point(0.4, 0.7))
    assert point_in_triangle(Point(0.4, 0.4), t, 1e-9) and not point_in_triangle(Point(0.8, 0.8), t, 1e-9)
    assert point_in_triangle(Point(0.1, 0.1), t, 1e-9)  # on a corner


def test_random_against_the_oracle():
    import oracle
    from numba_celltree_b200 import algorithms

    rng = np.random.default_rng(77)
    n = 200_000
    a, b = rng.uniform(-1, 3, (n, 2)), rng.uniform(-1, 3, (n, 2))
    lo = rng.uniform(-1, 2, (n, 2))
    boxes = np.column_stack((lo[:, 0], lo[:, 0] + rng.uniform(0, 2, n), lo[:, 1], lo[:, 1] + rng.uniform(0, 2, n)))
    for mine, theirs in (
        (algorithms.liang_barsky_line_box_clip, oracle.liang_barsky_line_box_clip),
        (algorithms.cohen_sutherland_line_box_clip, oracle.cohen_sutherland_line_box_clip_batch),
    ):
        got, want = mine(a, b, boxes), theirs(a, b, boxes)
        assert all(same(x, y) for x, y in zip(got, want))
    assert algorithms.liang_barsky_line_box_clip(a[:0], b[:0], boxes[:0])[0].shape == (0,)
