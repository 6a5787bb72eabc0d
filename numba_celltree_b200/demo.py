"""
Mesh helpers of the reference's ``demo`` module that do not draw (demo.py:11-43, 91-146): ``close_polygons``, ``edges``,
``generate_disk`` and ``example_1d_network``.  Pure NumPy / SciPy host code.  The plotting helpers (``plot_edges``,
``plot_boxes``: matplotlib) are out of scope.
"""

from __future__ import annotations

import numpy as np

from numba_celltree_b200.constants import IntArray, IntDType
from numba_celltree_b200.synthetic import generate_disk  # noqa: F401  (demo.py:91-123; triangulated with scipy)


def close_polygons(face_node_connectivity: IntArray, fill_value: int) -> IntArray:
    """Rows with one more column, in which every fill value -- and the new last column -- holds the row's first node,
    so that each row walks once around its polygon and ends where it started (demo.py:11-21)."""
    faces = np.asarray(face_node_connectivity)
    first = faces[:, :1]
    closed = np.concatenate((faces, first), axis=1).astype(IntDType)
    return np.where(closed == fill_value, first, closed)


def edges(face_node_connectivity: IntArray, fill_value: int) -> IntArray:
    """Unique edges ``(n_edge, 2)`` of a face array, each with its lower node first, in lexicographic order (demo.py:24-43)."""
    faces = np.atleast_2d(face_node_connectivity)
    n, m = faces.shape
    closed = close_polygons(faces, fill_value)
    pairs = np.stack((closed[:, :-1], closed[:, 1:]), axis=2).reshape(n * m, 2)
    pairs = np.sort(pairs, axis=1)
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]  # the sides a fill value collapsed to a point
    return np.unique(pairs, axis=0)


def example_1d_network():
    """The ten-vertex, nine-edge branching network of the reference's examples (demo.py:126-146)."""
    vertices = np.array(
        [[0.0, 0.0], [0.25, 1.0], [1.25, 2.0], [1.5, 2.5], [2.5, 3.25], [2.5, 2.5], [2.75, 3.75], [3.0, 2.0], [0.25, 1.75], [0.5, 2.25]]
    )
    network = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [3, 5], [4, 6], [5, 7], [1, 8], [8, 9]], dtype=np.int32)
    return vertices, network
