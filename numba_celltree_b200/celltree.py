"""
CellTree2d: the reference's public class (celltree.py:37-351) over the sm_100a library.

Same constructor, method names, argument meaning, return dtypes/shapes/order and error behaviour.  Every
array-level Numba kernel the reference's methods call is replaced by one C-ABI call (include/celltree_b200.h).
Inputs may also be float64 CUDA tensors (torch): results then come back as CUDA tensors and nothing crosses PCIe.
"""

from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from numba_celltree_b200 import _lib
from numba_celltree_b200.cast import cast_bboxes, cast_edges, cast_faces, cast_vertices, check_faces_shape
from numba_celltree_b200.celltree_base import CellTree2dBase, DeviceTree, _aligned, _is_cuda_tensor, _ptr
from numba_celltree_b200.constants import FloatArray, IntArray, IntDType


class CellTree2d(CellTree2dBase):
    """
    Construct a cell tree from 2D vertices and a faces indexing array.

    Parameters
    ----------
    vertices: ndarray of floats with shape ``(n_point, 2)``
    faces: ndarray of integers with shape ``(n_face, n_max_vert)``, padded with ``fill_value``
    fill_value: int
    n_buckets: int, optional, default: 4 (>= 2)
    cells_per_leaf: int, optional, default: 2 (>= 1)
    """

    _KIND = _lib.CT_KIND_FACES

    def __init__(
        self,
        vertices: FloatArray,
        faces: IntArray,
        fill_value: int,
        n_buckets: int = 4,
        cells_per_leaf: int = 2,
    ):
        if n_buckets < 2:
            raise ValueError("n_buckets must be >= 2")
        if cells_per_leaf < 1:
            raise ValueError("cells_per_leaf must be >= 1")
        vertices = cast_vertices(vertices, copy=True)
        faces = cast_faces(faces, fill_value)
        handle = ctypes.c_void_p()
        # counter_clockwise + build_face_bboxes + creation.initialize + bbox_tree, all on the device
        _lib.check(
            _lib.load().ct_tree_create(
                vertices.ctypes.data, vertices.shape[0], faces.ctypes.data, faces.shape[0], faces.shape[1],
                _lib.CT_KIND_FACES, int(n_buckets), int(cells_per_leaf), 0.0, _lib.CT_MEM_HOST, ctypes.byref(handle),
            )
        )  # fmt: skip
        self._tree = DeviceTree(handle.value)
        self.vertices = vertices
        self.n_buckets = n_buckets
        self.cells_per_leaf = cells_per_leaf

    @classmethod
    def from_arrays(cls, vertices, faces, nodes, bb_indices, bb_coords, cells_per_leaf: int = 2, n_buckets: int = 4):
        """Upload a tree whose arrays were built elsewhere (``faces`` already counter-clockwise, -1 filled)."""
        faces = cast_faces(faces, -1)
        return cls._from_arrays(vertices, faces, nodes, bb_indices, bb_coords, cells_per_leaf, n_buckets, _lib.CT_KIND_FACES)

    @property
    def faces(self):
        """The faces after counter_clockwise() (the reference rewrites its own copy: celltree.py:75-76)."""
        return self._download("elements")

    def _elements(self):
        return self.faces

    def locate_points(self, points: FloatArray, tolerance: Optional[float] = None, *, out=None) -> IntArray:
        """
        Find the index of a face that contains a point (-1 if none); points within ``tolerance`` of an edge
        count as inside.  ``tolerance=None``: 1e-12 x the largest bounding-box diagonal (at least 1e-15).
        ``out`` (extension): a preallocated intp result array, e.g. in pinned host memory.
        """
        return self._locate_points(points, tolerance, with_weights=False, out=out)

    def compute_barycentric_weights(
        self, points: FloatArray, tolerance: Optional[float] = None
    ) -> Tuple[IntArray, FloatArray]:
        """
        Face index and barycentric weights ``(n_point, n_max_vert)`` of every point (area-ratio weights for
        triangles, Wachspress coordinates otherwise); zero rows for points outside the mesh.
        """
        return self._locate_points(points, tolerance, with_weights=True)

    def _boxes(self, bbox_coords, with_area: bool):
        device = None
        if _is_cuda_tensor(bbox_coords):
            if bbox_coords.dim() != 2 or bbox_coords.shape[1] != 4 or str(bbox_coords.dtype) != "torch.float64":
                raise ValueError("bbox_coords must have shape (n_box, 4)")
            bbox_coords = _aligned(bbox_coords)
            device = bbox_coords.device
        else:
            bbox_coords = cast_bboxes(bbox_coords)
        return self._variable(
            "ct_locate_boxes", _ptr(bbox_coords), bbox_coords.shape[0], int(with_area), tensors=(bbox_coords,), device=device
        )

    def locate_boxes(self, bbox_coords: FloatArray) -> Tuple[IntArray, IntArray]:
        """Pairs (box index, face index) whose bounding boxes overlap; rows are ``(xmin, xmax, ymin, ymax)``."""
        i, j, _ = self._boxes(bbox_coords, with_area=False)
        return i, j

    def intersect_boxes(self, bbox_coords: FloatArray) -> Tuple[IntArray, IntArray, FloatArray]:
        """Pairs (box index, face index) with a positive area of intersection, and that area."""
        return self._boxes(bbox_coords, with_area=True)

    def _faces(self, vertices, faces, fill_value: int, write_back: bool, with_area: bool, device=None):
        return self._variable(
            "ct_locate_faces", _ptr(vertices), vertices.shape[0], _ptr(faces), faces.shape[0], faces.shape[1], int(fill_value),
            int(write_back), int(with_area), tensors=(vertices, faces), device=device,
        )  # fmt: skip

    @staticmethod
    def _device_mesh(vertices, faces):
        """A query mesh given as CUDA tensors (float64 ``(n, 2)``, int64 ``(n_face, n_max_vert)``): checked, contiguous."""
        if not (_is_cuda_tensor(vertices) and _is_cuda_tensor(faces)):
            raise ValueError("vertices and faces must both be CUDA tensors or both be host arrays")
        if vertices.dim() != 2 or vertices.shape[1] != 2 or str(vertices.dtype) != "torch.float64":
            raise ValueError("vertices must have shape (n_points, 2)")
        if str(faces.dtype) != "torch.int64":
            raise ValueError("faces must be an int64 tensor of shape (n_face, n_max_vert)")
        check_faces_shape(faces)
        if vertices.device != faces.device:
            raise ValueError("vertices and faces must be on the same device")
        return _aligned(vertices), _aligned(faces)

    def locate_faces(self, vertices: FloatArray, faces: IntArray) -> Tuple[IntArray, IntArray]:
        """
        Pairs (face index, tree face index) that overlap according to the separating axis theorem.
        As in the reference (celltree.py:212) ``faces`` is made counter-clockwise IN PLACE when it is a
        contiguous intp array (or a contiguous int64 CUDA tensor; results are then CUDA tensors).
        """
        if _is_cuda_tensor(vertices) or _is_cuda_tensor(faces):
            vertices_c, faces_c = self._device_mesh(vertices, faces)
            i, j, _ = self._faces(vertices_c, faces_c, -1, write_back=True, with_area=False, device=faces_c.device)
            if faces_c is not faces:
                faces.copy_(faces_c)
            return i, j
        vertices_c = np.ascontiguousarray(vertices, dtype=np.float64)
        faces_c = np.ascontiguousarray(faces, dtype=IntDType)
        if vertices_c.ndim != 2 or vertices_c.shape[1] != 2:
            raise ValueError("vertices must have shape (n_points, 2)")
        if faces_c.ndim != 2:
            raise ValueError("faces must have shape (n_face, n_max_vert)")
        i, j, _ = self._faces(vertices_c, faces_c, -1, write_back=True, with_area=False)
        if faces_c is not faces and isinstance(faces, np.ndarray) and faces.shape == faces_c.shape:
            faces[...] = faces_c
        return i, j

    def intersect_faces(
        self, vertices: FloatArray, faces: IntArray, fill_value: int
    ) -> Tuple[IntArray, IntArray, FloatArray]:
        """Pairs (face index, tree face index) with a positive area of intersection, and that area.
        With CUDA tensors as the query mesh the three results are CUDA tensors: the (i, j, area) triplets of the
        regridding weights stay on the device."""
        if _is_cuda_tensor(vertices) or _is_cuda_tensor(faces):
            vertices, faces = self._device_mesh(vertices, faces)
            return self._faces(vertices, faces, fill_value, write_back=False, with_area=True, device=faces.device)
        vertices = cast_vertices(vertices)
        if isinstance(faces, np.ndarray) and faces.dtype == IntDType and faces.flags.c_contiguous:
            # cast_faces' checks without its copy: the device works on its own copy of the faces and treats
            # fill_value as padding, so the caller's array is neither modified nor duplicated on the host
            check_faces_shape(faces)
        else:
            faces = cast_faces(faces, fill_value)
            fill_value = -1
        return self._faces(vertices, faces, fill_value, write_back=False, with_area=True)

    def intersect_edges(self, edge_coords: FloatArray) -> Tuple[IntArray, IntArray, FloatArray]:
        """
        Pairs (edge index, face index) and the part ``((x0, y0), (x1, y1))`` of each edge inside that face,
        ordered per edge by distance along the edge.
        """
        device = None
        if _is_cuda_tensor(edge_coords):
            if edge_coords.dim() != 3 or tuple(edge_coords.shape[1:]) != (2, 2) or str(edge_coords.dtype) != "torch.float64":
                raise ValueError("edges must have shape (n_edge, 2, 2)")
            edge_coords = _aligned(edge_coords)
            device = edge_coords.device
        else:
            edge_coords = cast_edges(edge_coords)
        return self._variable(
            "ct_intersect_edges", _ptr(edge_coords), edge_coords.shape[0], tensors=(edge_coords,), payload_shape=(2, 2), device=device
        )
