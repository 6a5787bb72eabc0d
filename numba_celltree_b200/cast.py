"""
Boundary hygiene: every array that crosses into the C-ABI is made C-contiguous with the library's dtypes, and the
ValueErrors users of the reference rely on are raised here (reference: cast.py:14-54; its tests pin the messages'
conditions, tests/test_celltree.py:134-164).
"""

import numpy as np

from numba_celltree_b200.constants import FILL_VALUE, MAX_N_VERTEX, FloatDType, IntDType

# name -> (dtype, number of dimensions, required trailing shape, message)
_LAYOUTS = {
    "vertices": (FloatDType, 2, (2,), "vertices must have shape (n_points, 2)"),
    "bboxes": (FloatDType, 2, (4,), "bbox_coords must have shape (n_box, 4)"),
    "edges": (FloatDType, 3, (2, 2), "edges must have shape (n_edge, 2, 2)"),
}


def _conform(values, layout: str, copy: bool = False):
    dtype, ndim, trailing, message = _LAYOUTS[layout]
    array = values.astype(dtype, copy=copy) if isinstance(values, np.ndarray) else np.asarray(values, dtype=dtype)
    if array.ndim != ndim or array.shape[1:] != trailing:
        raise ValueError(message)
    return np.ascontiguousarray(array)


def cast_vertices(vertices, copy: bool = False):
    return _conform(vertices, "vertices", copy)


def cast_bboxes(bbox_coords):
    return _conform(bbox_coords, "bboxes")


def cast_edges(edges):
    return _conform(edges, "edges")


def check_faces_shape(faces) -> None:
    """The two conditions cast_faces rejects, for callers that hand their own intp array straight to the device."""
    if faces.ndim != 2:
        raise ValueError("faces must have shape (n_face, n_max_vert)")
    widest = faces.shape[1]
    if widest > MAX_N_VERTEX:
        raise ValueError(
            f"faces contains up to {widest} vertices for a single face; at most {MAX_N_VERTEX} vertices per face are supported."
        )


def cast_faces(faces, fill_value: int):
    """A private intp copy of `faces` with `fill_value` rewritten to -1 (the copy is what counter_clockwise may reorder)."""
    own = np.array(faces, dtype=IntDType, order="C", copy=True)
    check_faces_shape(own)
    if fill_value != FILL_VALUE:
        np.putmask(own, own == fill_value, FILL_VALUE)
    return own
