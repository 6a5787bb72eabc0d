"""
EdgeCellTree2d: the reference's public class for 1-D networks (edge_celltree.py:25-146) over the sm_100a library.
"""

from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from numba_celltree_b200 import _lib
from numba_celltree_b200.cast import cast_edges, cast_vertices
from numba_celltree_b200.celltree_base import CellTree2dBase, DeviceTree, _aligned, _is_cuda_tensor, _ptr
from numba_celltree_b200.constants import MIN_TOLERANCE, TOLERANCE_FACTOR, FloatArray, IntArray, IntDType


class EdgeCellTree2d(CellTree2dBase):
    """
    Construct a cell tree from 2D vertices and an edges indexing array.

    Parameters
    ----------
    vertices: ndarray of floats with shape ``(n_point, 2)``
    edges: ndarray of integers with shape ``(n_edge, 2)``
    n_buckets: int, optional, default: 4 (>= 2)
    cells_per_leaf: int, optional, default: 2 (>= 1)
    """

    _KIND = _lib.CT_KIND_EDGES

    def __init__(self, vertices: FloatArray, edges: IntArray, n_buckets: int = 4, cells_per_leaf: int = 2):
        if n_buckets < 2:
            raise ValueError("n_buckets must be >= 2")
        if cells_per_leaf < 1:
            raise ValueError("cells_per_leaf must be >= 1")
        vertices = cast_vertices(vertices, copy=True)
        # padding of the edge bounding boxes (axis-aligned edges have no width): edge_celltree.py:59-64
        x, y = vertices.T
        dx = x.max() - x.min()
        dy = y.max() - y.min()
        global_tolerance = max(MIN_TOLERANCE, TOLERANCE_FACTOR * max(dx, dy))
        edges_c = np.ascontiguousarray(edges, dtype=IntDType)
        if edges_c.ndim != 2 or edges_c.shape[1] != 2:
            raise ValueError("edges must have shape (n_edge, 2)")
        handle = ctypes.c_void_p()
        _lib.check(
            _lib.load().ct_tree_create(
                vertices.ctypes.data, vertices.shape[0], edges_c.ctypes.data, edges_c.shape[0], 2,
                _lib.CT_KIND_EDGES, int(n_buckets), int(cells_per_leaf), float(global_tolerance), _lib.CT_MEM_HOST,
                ctypes.byref(handle),
            )
        )  # fmt: skip
        self._tree = DeviceTree(handle.value)
        self.vertices = vertices
        self.edges = edges  # kept as given, like the reference (edge_celltree.py:68)
        self.n_buckets = n_buckets
        self.cells_per_leaf = cells_per_leaf

    def _elements(self):
        return self.edges

    def locate_points(self, points: FloatArray, tolerance: Optional[float] = None) -> IntArray:
        """Index of an edge within ``tolerance`` of each point, -1 if none."""
        return self._locate_points(points, tolerance, with_weights=False)

    def intersect_edges(self, edge_coords: FloatArray) -> Tuple[IntArray, IntArray, FloatArray]:
        """Pairs (edge index, tree edge index) and the intersection point, ordered per edge along the edge.
        A float64 CUDA tensor of shape ``(n_edge, 2, 2)`` is read in place and the results are CUDA tensors."""
        device = None
        if _is_cuda_tensor(edge_coords):
            if edge_coords.dim() != 3 or tuple(edge_coords.shape[1:]) != (2, 2) or str(edge_coords.dtype) != "torch.float64":
                raise ValueError("edges must have shape (n_edge, 2, 2)")
            edge_coords = _aligned(edge_coords)
            device = edge_coords.device
        else:
            edge_coords = cast_edges(edge_coords)
        i, j, xy = self._variable(
            "ct_intersect_edges", _ptr(edge_coords), edge_coords.shape[0], tensors=(edge_coords,), payload_shape=(2, 2), device=device
        )
        if device is not None:
            return i, j, xy[:, 0].contiguous()
        return i, j, np.ascontiguousarray(xy[:, 0])
