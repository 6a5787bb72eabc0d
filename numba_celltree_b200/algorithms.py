"""
The geometry helpers the reference exports beside its query API (numba_celltree.algorithms.__all__, plus
point_in_polygon / point_in_triangle / points_in_triangles of geometry_utils.py), over the sm_100a library.

The reference's functions are scalar ``@njit`` functions on ``Point`` / ``Box`` tuples.  Here every function takes
either those scalars (same return value as the reference: ``(bool, Point, Point)`` or ``bool``) or arrays of n
inputs (returns arrays): one C-ABI call, one thread per input, the same fp64 device functions as the query kernels.
There is no CPU path.
"""

from __future__ import annotations

from typing import NamedTuple

import numpy as np

from numba_celltree_b200 import _lib
from numba_celltree_b200.constants import FloatDType, IntDType

__all__ = (
    "Point", "Box", "Triangle",
    "liang_barsky_line_box_clip", "cohen_sutherland_line_box_clip", "cyrus_beck_line_polygon_clip",
    "point_in_polygon", "points_in_polygon", "point_in_triangle", "points_in_triangles",
)  # fmt: skip


class Point(NamedTuple):  # constants.py:36-38
    x: float
    y: float


class Box(NamedTuple):  # constants.py:51-55
    xmin: float
    xmax: float
    ymin: float
    ymax: float


class Triangle(NamedTuple):  # constants.py:58-61
    a: Point
    b: Point
    c: Point


def _f64(values, shape, message):
    array = np.ascontiguousarray(values, dtype=FloatDType)
    if array.ndim == len(shape) - 1:  # a single Point / Box
        array = array.reshape((1,) + array.shape)
    if array.ndim != len(shape) or array.shape[1:] != shape[1:]:
        raise ValueError(message)
    return array


def _is_scalar(a) -> bool:
    return np.ndim(a) == 1


def _clip_result(scalar, hit, c, d):
    if scalar:
        return bool(hit[0]), Point(float(c[0, 0]), float(c[0, 1])), Point(float(d[0, 0]), float(d[0, 1]))
    return hit.astype(bool), c, d


def _line_box_clip(entry: str, a, b, box):
    scalar = _is_scalar(a) and _is_scalar(b) and _is_scalar(box)
    a = _f64(a, (None, 2), "a must be a Point or have shape (n, 2)")
    b = _f64(b, (None, 2), "b must be a Point or have shape (n, 2)")
    boxes = _f64(box, (None, 4), "box must be a Box or have shape (n, 4)")
    n = len(a)
    if len(b) != n or len(boxes) not in (1, n):
        raise ValueError("a and b must have the same length; box must be one Box or one per segment")
    hit = np.empty(n, dtype=np.uint8)
    c = np.empty((n, 2), dtype=FloatDType)
    d = np.empty((n, 2), dtype=FloatDType)
    _lib.check(
        getattr(_lib.load(), entry)(
            a.ctypes.data, b.ctypes.data, boxes.ctypes.data, len(boxes), n, hit.ctypes.data, c.ctypes.data, d.ctypes.data, _lib.CT_MEM_HOST
        )
    )
    return _clip_result(scalar, hit, c, d)


def liang_barsky_line_box_clip(a, b, box):
    """Liang-Barsky clip of segment a -> b against ``box`` (algorithms/liang_barsky.py:10-64)."""
    return _line_box_clip("ct_liang_barsky_line_box_clip", a, b, box)


def cohen_sutherland_line_box_clip(a, b, box):
    """Cohen-Sutherland clip of segment a -> b against ``box`` (algorithms/cohen_sutherland.py:36-101)."""
    return _line_box_clip("ct_cohen_sutherland_line_box_clip", a, b, box)


def cyrus_beck_line_polygon_clip(a, b, poly, tolerance: float):
    """Cyrus-Beck clip of segment(s) a -> b against one counter-clockwise convex polygon (algorithms/cyrus_beck.py:143-241)."""
    scalar = _is_scalar(a) and _is_scalar(b)
    a = _f64(a, (None, 2), "a must be a Point or have shape (n, 2)")
    b = _f64(b, (None, 2), "b must be a Point or have shape (n, 2)")
    poly = _f64(poly, (None, 2), "poly must have shape (n_vertex, 2)")
    n = len(a)
    if len(b) != n:
        raise ValueError("a and b must have the same length")
    hit = np.empty(n, dtype=np.uint8)
    c = np.empty((n, 2), dtype=FloatDType)
    d = np.empty((n, 2), dtype=FloatDType)
    _lib.check(
        _lib.load().ct_cyrus_beck_line_polygon_clip(
            a.ctypes.data, b.ctypes.data, n, poly.ctypes.data, len(poly), float(tolerance), hit.ctypes.data, c.ctypes.data,
            d.ctypes.data, _lib.CT_MEM_HOST,
        )
    )  # fmt: skip
    return _clip_result(scalar, hit, c, d)


def points_in_polygon(points, poly):
    """Crossing-number test of every point against one polygon, no tolerance (geometry_utils.py:98-147)."""
    points = _f64(points, (None, 2), "points must have shape (n, 2)")
    poly = _f64(poly, (None, 2), "poly must have shape (n_vertex, 2)")
    inside = np.empty(len(points), dtype=np.uint8)
    _lib.check(_lib.load().ct_points_in_polygon(points.ctypes.data, len(points), poly.ctypes.data, len(poly), inside.ctypes.data, _lib.CT_MEM_HOST))
    return inside.astype(bool)


def point_in_polygon(p, poly) -> bool:
    """``point_in_polygon(Point, poly) -> bool`` as in the reference (geometry_utils.py:98-147)."""
    return bool(points_in_polygon(np.asarray(p, dtype=FloatDType).reshape(1, 2), poly)[0])


def points_in_triangles(points, face_indices, faces, vertices, tolerance: float):
    """For every point, whether it lies in (or within ``tolerance`` of an edge of) the triangle ``faces[face_indices[i]]``
    (geometry_utils.py:273-289)."""
    points = _f64(points, (None, 2), "points must have shape (n, 2)")
    face_indices = np.ascontiguousarray(face_indices, dtype=IntDType)
    faces = np.ascontiguousarray(faces, dtype=IntDType)
    vertices = _f64(vertices, (None, 2), "vertices must have shape (n_vertex, 2)")
    if face_indices.shape != (len(points),) or faces.ndim != 2 or faces.shape[1] < 3:
        raise ValueError("face_indices must have shape (n,), faces (n_face, >= 3)")
    inside = np.empty(len(points), dtype=np.uint8)
    _lib.check(
        _lib.load().ct_points_in_triangles(
            points.ctypes.data, face_indices.ctypes.data, len(points), faces.ctypes.data, faces.shape[0], faces.shape[1],
            vertices.ctypes.data, len(vertices), float(tolerance), inside.ctypes.data, _lib.CT_MEM_HOST,
        )
    )  # fmt: skip
    return inside.astype(bool)


def point_in_triangle(p, t, tolerance: float) -> bool:
    """``point_in_triangle(Point, Triangle, tolerance) -> bool`` as in the reference (geometry_utils.py:241-270)."""
    vertices = np.array([t[0], t[1], t[2]], dtype=FloatDType)
    return bool(points_in_triangles(np.asarray(p, dtype=FloatDType).reshape(1, 2), [0], [[0, 1, 2]], vertices, tolerance)[0])
