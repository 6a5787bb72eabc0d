"""
Boundary hygiene: dtype / shape normalisation and the ValueErrors of the reference's cast.py:14-54
(the error conventions are part of the drop-in contract; the reference's tests pin them).
"""

import numpy as np

from numba_celltree_b200.constants import FILL_VALUE, MAX_N_VERTEX, FloatDType, IntDType


def _as_array(values, dtype, copy):
    if isinstance(values, np.ndarray):
        return values.astype(dtype, copy=copy)
    return np.ascontiguousarray(values, dtype=dtype)


def cast_vertices(vertices, copy: bool = False):
    vertices = _as_array(vertices, FloatDType, copy)
    if vertices.ndim != 2 or vertices.shape[1] != 2:
        raise ValueError("vertices must have shape (n_points, 2)")
    return np.ascontiguousarray(vertices)


def check_faces_shape(faces) -> None:
    if faces.ndim != 2:
        raise ValueError("faces must have shape (n_face, n_max_vert)")
    n_max_vert = faces.shape[1]
    if n_max_vert > MAX_N_VERTEX:
        raise ValueError(
            f"faces contains up to {n_max_vert} vertices for a single face. "
            f"A maximum of {MAX_N_VERTEX} vertices per face is supported."
        )


def cast_faces(faces, fill_value: int):
    faces = _as_array(faces, IntDType, True)
    check_faces_shape(faces)
    if fill_value != FILL_VALUE:
        faces[faces == fill_value] = FILL_VALUE
    return np.ascontiguousarray(faces)


def cast_bboxes(bbox_coords):
    bbox_coords = np.ascontiguousarray(bbox_coords, dtype=FloatDType)
    if bbox_coords.ndim != 2 or bbox_coords.shape[1] != 4:
        raise ValueError("bbox_coords must have shape (n_box, 4)")
    return bbox_coords


def cast_edges(edges):
    edges = np.ascontiguousarray(edges, dtype=FloatDType)
    if edges.ndim != 3 or edges.shape[1] != 2 or edges.shape[2] != 2:
        raise ValueError("edges must have shape (n_edge, 2, 2)")
    return edges
