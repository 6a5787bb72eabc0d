"""The CPU oracle against the reference's own known-answer tests (restated in tests/_reference_kats.py)."""

import numpy as np

import oracle
from tests import _reference_kats as K


def test_init_and_casting():
    K.check_init_and_casting(oracle.CellTree2d)


def test_errors():
    K.check_errors(oracle.CellTree2d)


def test_point_lookups():
    K.check_point_lookups(oracle.CellTree2d)


def test_box_and_edge_lookup():
    K.check_box_and_edge_lookup(oracle.CellTree2d)


def test_example_material():
    K.check_example_material(oracle.CellTree2d)


def test_barycentric():
    K.check_barycentric(oracle.CellTree2d)


def test_locate_point_on_edge():
    K.check_locate_point_on_edge(oracle.CellTree2d)


def test_edge_tree():
    K.check_edge_tree(oracle.EdgeCellTree2d, oracle.CellTreeData)


# ---- scalar known answers: tests/test_algorithms/test_line_box_clip.py:42-128 --------------------------------
BOX = (0.0, 2.0, 0.0, 2.0)
POLY = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 2.0], [0.0, 2.0]])


def _clippers():
    return [
        lambda a, b: oracle.cohen_sutherland_line_box_clip(a, b, BOX),
        lambda a, b: oracle.cyrus_beck_line_polygon_clip(a, b, POLY, 1e-9),
    ]


def test_line_box_clip_known_answers():
    table = [
        ((-1.0, 0.0), (2.0, 3.0), True, (0.0, 1.0), (1.0, 2.0)),
        ((0.0, -0.1), (0.0, -0.1), False, None, None),
        ((-1.0, 1.0), (3.0, 1.0), True, (0.0, 1.0), (2.0, 1.0)),
        ((1.0, -3.0), (1.0, 3.0), True, (1.0, 0.0), (1.0, 2.0)),
        ((1.0, -3.0), (1.0, 1.0), True, (1.0, 0.0), (1.0, 1.0)),
        ((1.0, 1.0), (1.0, 3.0), True, (1.0, 1.0), (1.0, 2.0)),
        ((-1.0, 3.0), (3.0, 3.0), False, None, None),
        ((-1.0, 1.0), (1.0, 1.0), True, (0.0, 1.0), (1.0, 1.0)),
        ((0.5, 0.5), (1.5, 1.5), True, (0.5, 0.5), (1.5, 1.5)),
        ((-1.5, 0.0), (-0.5, 1.0), False, None, None),
        ((2.5, 0.0), (3.5, 1.0), False, None, None),
    ]
    for clip in _clippers():
        for a, b, expected, c, d in table:
            ok, cc, dd = clip(a, b)
            assert ok == expected, (a, b)
            if expected:
                assert np.allclose(cc, c) and np.allclose(dd, d), (a, b, cc, dd)
            else:
                assert np.isnan(cc).all() and np.isnan(dd).all()


def test_clip_area_known_answers():
    # tests/test_algorithms/test_sutherland_hodgman.py: unit-square style checks
    square = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    assert oracle.polygon_polygon_clip_area(square, square + 0.5) == 0.25
    assert oracle.polygon_polygon_clip_area(square, square + [1.0, 0.0]) == 0.0  # sharing an edge
    assert oracle.polygon_polygon_clip_area(square, square + 2.0) == 0.0
    tri = np.array([[0.0, 0.0], [2.0, 0.0], [0.0, 2.0]])
    assert oracle.polygon_polygon_clip_area(tri, square) == 1.0
    # repeated vertices are skipped (:113-139)
    tri_rep = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 0.0], [0.0, 2.0]])
    assert oracle.polygon_polygon_clip_area(tri_rep, square) == 1.0


def test_separating_axes_known_answers():
    # tests/test_algorithms/test_separating_axis.py:80-100: touching is not intersecting; both directions needed
    a = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    b = a + [1.0, 0.0]
    assert not (oracle.separating_axes(a, b) and oracle.separating_axes(b, a))
    c = a + [0.5, 0.5]
    assert oracle.separating_axes(a, c) and oracle.separating_axes(c, a)
    # hanging / repeated nodes (:185-223)
    d = np.array([[0.0, 0.0], [0.5, 0.0], [1.0, 0.0], [1.0, 1.0], [1.0, 1.0], [0.0, 1.0]])
    assert oracle.separating_axes(d, c) and oracle.separating_axes(c, d)


def test_lines_intersect_known_answers():
    # tests/test_geometry_utils.py:354-419
    ok, x, y = oracle.lines_intersect((0.0, 0.0), (2.0, 2.0), (0.0, 2.0), (2.0, 0.0))
    assert ok and (x, y) == (1.0, 1.0)
    ok, x, y = oracle.lines_intersect((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0))  # parallel
    assert not ok and np.isnan(x) and np.isnan(y)
    ok, x, y = oracle.lines_intersect((0.0, 0.0), (2.0, 0.0), (1.0, 0.0), (3.0, 0.0))  # collinear overlap -> midpoint
    assert ok and (x, y) == (1.5, 0.0)
    ok, x, y = oracle.lines_intersect((0.0, 0.0), (0.0, 0.0), (1.0, 0.0), (3.0, 0.0))  # no length
    assert not ok
