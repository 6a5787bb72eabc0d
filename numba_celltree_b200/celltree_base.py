"""
Shared host-side machinery of CellTree2d and EdgeCellTree2d: the device tree handle, lazily mirrored
attributes, and the marshalling of NumPy arrays / CUDA tensors into the C-ABI.

Mirrors the reference's celltree_base.py:55-114 (CellTree2dBase) and the attribute set of
celltree.py:80-97 / edge_celltree.py:66-83.
"""

from __future__ import annotations

import abc
import ctypes
from typing import Optional

import numpy as np

from numba_celltree_b200 import _lib
from numba_celltree_b200.constants import CellTreeData, FloatDType, IntDType, NodeDType


def _is_cuda_tensor(x) -> bool:
    return bool(getattr(x, "is_cuda", False)) and hasattr(x, "data_ptr")


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class DeviceTree:
    """Owner of one ``ct_tree*``."""

    def __init__(self, handle: int):
        self.handle = ctypes.c_void_p(handle)
        info = _lib.TreeInfo()
        _lib.check(_lib.load().ct_tree_get_info(self.handle, ctypes.byref(info)))
        self.info = info

    def close(self):
        if self.handle is not None and self.handle.value:
            _lib.load().ct_tree_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CellTree2dBase(abc.ABC):
    _tree: DeviceTree

    # ---- mirrors of the device arrays (downloaded on first access, then cached and writable) -------------
    def _download(self, name: str):
        cache = self.__dict__.setdefault("_mirrors", {})
        if name not in cache:
            info = self._tree.info
            n, m = info.n_elem, info.n_max_vert
            shapes = {
                "nodes": ((info.n_nodes,), NodeDType),
                "bb_indices": ((n,), IntDType),
                "bb_coords": ((n, 4), FloatDType),
                "elements": ((n, m), IntDType),
                "bb_distances": ((n, 3), FloatDType),
            }
            shape, dtype = shapes[name]
            out = np.empty(shape, dtype=dtype)
            order = ["nodes", "bb_indices", "bb_coords", "elements", "bb_distances"]
            args = [out.ctypes.data if key == name else None for key in order]
            _lib.check(_lib.load().ct_tree_download(self._tree.handle, *args, _lib.CT_MEM_HOST))
            cache[name] = out
        return cache[name]

    @property
    def nodes(self):
        return self._download("nodes")

    @property
    def bb_indices(self):
        return self._download("bb_indices")

    @property
    def bb_coords(self):
        return self._download("bb_coords")

    @property
    def bb_distances(self):
        return self._download("bb_distances")

    @property
    def bbox(self):
        if "_bbox" not in self.__dict__:
            self._bbox = np.array(list(self._tree.info.bbox), dtype=FloatDType)
        return self._bbox

    @property
    def celltree_data(self) -> CellTreeData:
        return CellTreeData(
            self._elements(), self.vertices, self.nodes, self.bb_indices, self.bb_coords, self.bbox, self.cells_per_leaf
        )

    @abc.abstractmethod
    def _elements(self):
        pass

    @property
    def depth(self) -> int:
        """Number of node levels of the tree."""
        return int(self._tree.info.depth)

    @property
    def build_ms(self) -> float:
        """Device time of the tree construction (CUDA events)."""
        return float(self._tree.info.build_ms)

    def _default_tolerance(self) -> float:
        # default_tolerance(bb_distances[:, 2]) of celltree_base.py:51-52, computed once at build time
        return float(self._tree.info.default_tolerance)

    # ---- fixed-size queries ---------------------------------------------------------------------------------
    def _locate_points(self, points, tolerance: Optional[float], with_weights: bool, out=None):
        from numba_celltree_b200.cast import cast_vertices

        if tolerance is None:
            tolerance = self._default_tolerance()
        lib = _lib.load()
        m = int(self._tree.info.n_max_vert)
        if _is_cuda_tensor(points):
            import torch

            if points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] != 2:
                raise ValueError("points must be a float64 CUDA tensor of shape (n_points, 2)")
            points = points.contiguous()
            n = points.shape[0]
            out = torch.empty(n, dtype=torch.int64, device=points.device)
            weights = torch.empty((n, m), dtype=torch.float64, device=points.device) if with_weights else None
            mem = _lib.CT_MEM_DEVICE
        else:
            points = cast_vertices(points)
            n = points.shape[0]
            if out is None:
                out = _lib.result_array(n, IntDType)
            elif not (isinstance(out, np.ndarray) and out.dtype == IntDType and out.shape == (n,) and out.flags.c_contiguous):
                raise ValueError("out must be a C-contiguous intp array of shape (n_points,)")
            weights = _lib.result_array((n, m), FloatDType) if with_weights else None
            mem = _lib.CT_MEM_HOST
        _lib.check(lib.ct_locate_points(self._tree.handle, _ptr(points), n, float(tolerance), _ptr(out), _ptr(weights), mem))
        return (out, weights) if with_weights else out

    # ---- variable-length results --------------------------------------------------------------------------------
    @staticmethod
    def _fetch(handle: ctypes.c_void_p, payload_shape=None, device=None):
        lib = _lib.load()
        try:
            total = lib.ct_result_size(handle)
            width = lib.ct_result_payload_width(handle)
            if device is None:
                i = _lib.result_array(total, IntDType)
                j = _lib.result_array(total, IntDType)
                payload = _lib.result_array((total,) + tuple(payload_shape), FloatDType) if width else None
                mem = _lib.CT_MEM_HOST
            else:
                import torch

                i = torch.empty(total, dtype=torch.int64, device=device)
                j = torch.empty(total, dtype=torch.int64, device=device)
                payload = torch.empty((total,) + tuple(payload_shape), dtype=torch.float64, device=device) if width else None
                mem = _lib.CT_MEM_DEVICE
            _lib.check(lib.ct_result_fetch(handle, _ptr(i), _ptr(j), _ptr(payload), mem))
        finally:
            lib.ct_result_free(handle)
        return i, j, payload

    # ---- diagnostics (host side; serial O(n_nodes) in the reference too: query.py:568-664) -------------------
    @property
    def node_bounds(self):
        """Bounds (xmin, xmax, ymin, ymax) of every node: collect_node_bounds, query.py:568-621."""
        nodes = self.nodes
        bounds = np.empty((len(nodes), 4), dtype=FloatDType)
        bounds[0] = self.bbox
        child = nodes["child"]
        dim = nodes["dim"].astype(np.intp)
        # children have larger indices than their parent, so one ascending pass visits parents first
        for parent in np.flatnonzero(child != -1):
            left = child[parent]
            right = left + 1
            bounds[left] = bounds[parent]
            bounds[right] = bounds[parent]
            bounds[left, 2 * dim[parent] + 1] = nodes["Lmax"][parent]
            bounds[right, 2 * dim[parent]] = nodes["Rmin"][parent]
        return bounds

    def validate_node_bounds(self):
        """For every node, whether its children (or its cells' boxes) lie within its bounds: query.py:624-664."""
        nodes = self.nodes
        bounds = self.node_bounds
        bb_coords = self.bb_coords
        bb_indices = self.bb_indices
        valid = np.zeros(len(nodes), dtype=bool)

        def contained(a, b):
            return (a[..., 0] >= b[0]) & (a[..., 1] <= b[1]) & (a[..., 2] >= b[2]) & (a[..., 3] <= b[3])

        for index in range(len(nodes)):
            node = nodes[index]
            box = bounds[index]
            if node["child"] == -1:
                cells = bb_indices[node["ptr"] : node["ptr"] + node["size"]]
                valid[index] = bool(np.all(contained(bb_coords[cells], box)))
            else:
                left = node["child"]
                valid[index] = bool(contained(bounds[left], box) and contained(bounds[left + 1], box))
        return valid

    def to_dict_of_lists(self):
        """Children of every node as ``{index: [left, right] or []}``: celltree_base.py:80-114."""
        result = {}
        for index, left in enumerate(self.nodes["child"]):
            result[index] = [] if left == -1 else [int(left), int(left) + 1]
        return result
