"""
Boundary hygiene of the product's host side (numba_celltree_b200/cast.py), without a GPU: the leniency and the
ValueErrors that users of the reference rely on (reference cast.py:14-54; its tests tests/test_celltree.py:115-164,
316-330 pin them).
"""

import numpy as np
import pytest

from numba_celltree_b200.cast import cast_bboxes, cast_edges, cast_faces, cast_vertices, check_faces_shape
from numba_celltree_b200.constants import MAX_N_VERTEX, FloatDType, IntDType


def test_lists_and_narrow_dtypes_are_accepted():
    v = cast_vertices([[0, 0], [1, 0], [0, 1]])
    assert v.dtype == FloatDType and v.shape == (3, 2) and v.flags.c_contiguous
    v32 = cast_vertices(np.array([[0, 0], [1, 0.5]], dtype=np.float32))
    assert v32.dtype == FloatDType and v32[1, 1] == 0.5
    f = cast_faces(np.array([[0, 1, 2]], dtype=np.int32), -1)
    assert f.dtype == IntDType and f.flags.c_contiguous
    b = cast_bboxes([[0.0, 1.0, 0.0, 1.0]])
    assert b.shape == (1, 4) and b.dtype == FloatDType
    e = cast_edges([[[0.0, 0.0], [1.0, 1.0]]])
    assert e.shape == (1, 2, 2) and e.dtype == FloatDType


def test_non_contiguous_inputs_become_contiguous_without_changing_values():
    wide = np.arange(40, dtype=np.float64).reshape(10, 4)
    view = wide[:, ::2]
    assert not view.flags.c_contiguous
    v = cast_vertices(view)
    assert v.flags.c_contiguous and np.array_equal(v, view)
    boxes = np.asfortranarray(np.arange(20, dtype=np.float64).reshape(5, 4))
    b = cast_bboxes(boxes)
    assert b.flags.c_contiguous and np.array_equal(b, boxes)


def test_copy_semantics():
    v = np.zeros((4, 2))
    assert cast_vertices(v) is v or np.shares_memory(cast_vertices(v), v)  # no copy for a conforming array
    assert not np.shares_memory(cast_vertices(v, copy=True), v)  # the constructor's own copy (celltree.py:74)
    f = np.array([[0, 1, 2]], dtype=IntDType)
    assert not np.shares_memory(cast_faces(f, -1), f)  # cast.py:26: faces are always copied (counter_clockwise reorders them)


def test_fill_value_is_rewritten_to_minus_one():
    f = cast_faces(np.array([[0, 1, 2, -999], [0, 2, 3, 4]]), -999)
    assert f[0, 3] == -1 and f[1, 3] == 4
    same = cast_faces(np.array([[0, 1, 2, -1]]), -1)
    assert same[0, 3] == -1
    # a fill value that also is a valid index elsewhere: every occurrence is rewritten, as cast.py:38-39 does
    g = cast_faces(np.array([[0, 1, 2, 3], [3, 1, 2, 3]]), 3)
    assert (g == -1).sum() == 3


@pytest.mark.parametrize(
    "call,bad",
    [
        (cast_vertices, np.zeros((3, 3))),
        (cast_vertices, np.zeros(6)),
        (cast_bboxes, np.zeros((3, 2))),
        (cast_bboxes, np.zeros((3, 4, 1))),
        (cast_edges, np.zeros((3, 4))),
        (cast_edges, np.zeros((3, 2, 3))),
    ],
)
def test_wrong_shapes_raise_value_error(call, bad):
    with pytest.raises(ValueError):
        call(bad)


def test_faces_shape_errors():
    with pytest.raises(ValueError):
        cast_faces(np.zeros(5, dtype=int), -1)
    with pytest.raises(ValueError):
        cast_faces(np.zeros((2, MAX_N_VERTEX + 1), dtype=int), -1)
    check_faces_shape(np.zeros((2, MAX_N_VERTEX), dtype=int))  # the widest face allowed (constants.py:128)
    with pytest.raises(ValueError):
        check_faces_shape(np.zeros((2, 2, 2), dtype=int))
