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


class _OnTensorStream:
    """
    Binds a device-resident call to the caller's tensors: they must live on the tree's device (a pointer from another GPU
    would fault inside the kernels), and the library's work is enqueued on torch's CURRENT stream of that device, so that
    it is ordered after the kernels that produced the inputs and before whatever consumes the results -- also under
    ``torch.cuda.stream(side)`` -- and a temporary made by ``.contiguous()`` is not reused while a kernel still reads it.
    """

    def __init__(self, tree: "DeviceTree", *tensors):
        import torch

        device = int(tree.info.device)
        for t in tensors:
            if t is not None and t.device.index != device:
                raise ValueError(f"the tensor lives on {t.device}, the tree on cuda:{device}")
        self.stream = torch.cuda.current_stream(device).cuda_stream

    def __enter__(self):
        _lib.check(_lib.load().ct_set_stream(self.stream))
        return self

    def __exit__(self, *exc):
        return False


def _aligned(tensor):
    """A contiguous tensor whose first byte is 16-byte aligned (the kernels read points and boxes as 16-byte words; a
    view into the middle of a buffer may start on an odd 8-byte boundary)."""
    tensor = tensor.contiguous()
    return tensor if tensor.data_ptr() % 16 == 0 else tensor.clone()


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _checksum(array: np.ndarray):
    """Wrapping sum and xor over the bytes of a contiguous array taken as 64-bit words (any edit of a field changes them,
    short of a deliberate collision)."""
    raw = array.reshape(-1).view(np.uint8)
    whole = raw[: raw.size // 8 * 8].view(np.uint64)
    return (int(np.add.reduce(whole, dtype=np.uint64)), int(np.bitwise_xor.reduce(whole)) if whole.size else 0, raw[whole.size * 8 :].tobytes())


class DeviceTree:
    """Owner of one ``ct_tree*``."""

    def __init__(self, handle: int):
        self.handle = ctypes.c_void_p(handle)
        self.refresh_info()

    def refresh_info(self):
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
    _KIND: int  # _lib.CT_KIND_FACES or _lib.CT_KIND_EDGES, set by the two concrete classes

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
            if name == "nodes":
                self._nodes_checksum = _checksum(out)
        return cache[name]

    def _sync_nodes(self) -> None:
        """
        The reference's queries read ``tree.nodes`` on every call (query.py:73), so a caller who edits that array changes
        the answers (tests/test_celltree.py:606-618).  Here the array is a host mirror of the device tree: once it has
        been handed out, every query first compares a checksum of it (one pass over the array, about 0.1 ms per MB)
        and, if it was edited, sends it to the device, which derives its traversal structures again.
        """
        nodes = self.__dict__.get("_mirrors", {}).get("nodes")
        if nodes is None:
            return
        now = _checksum(nodes)
        if now != self._nodes_checksum:
            _lib.check(_lib.load().ct_tree_update_nodes(self._tree.handle, nodes.ctypes.data, len(nodes), _lib.CT_MEM_HOST))
            self._nodes_checksum = now
            self._tree.refresh_info()

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
        self._sync_nodes()
        lib = _lib.load()
        m = int(self._tree.info.n_max_vert)
        if _is_cuda_tensor(points):
            import torch

            if points.dtype != torch.float64 or points.dim() != 2 or points.shape[1] != 2:
                raise ValueError("points must be a float64 CUDA tensor of shape (n_points, 2)")
            points = _aligned(points)
            n = points.shape[0]
            with _OnTensorStream(self._tree, points):
                out = torch.empty(n, dtype=torch.int64, device=points.device)
                weights = torch.empty((n, m), dtype=torch.float64, device=points.device) if with_weights else None
                _lib.check(lib.ct_locate_points(self._tree.handle, _ptr(points), n, float(tolerance), _ptr(out), _ptr(weights), _lib.CT_MEM_DEVICE))
            return (out, weights) if with_weights else out
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
    def _variable(self, entry: str, *args, tensors=(), payload_shape=(), device=None):
        """One variable-length query: `entry(tree, *args, mem, &result)` followed by the fetch of the pairs, on the
        caller's stream when the inputs are CUDA tensors."""
        self._sync_nodes()
        lib = _lib.load()
        handle = ctypes.c_void_p()
        if device is None:
            _lib.check(getattr(lib, entry)(self._tree.handle, *args, _lib.CT_MEM_HOST, ctypes.byref(handle)))
            return self._fetch(handle, payload_shape=payload_shape)
        with _OnTensorStream(self._tree, *tensors):
            _lib.check(getattr(lib, entry)(self._tree.handle, *args, _lib.CT_MEM_DEVICE, ctypes.byref(handle)))
            return self._fetch(handle, payload_shape=payload_shape, device=device)

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

    # ---- serialisation (the reference has none, a 16.7 M-cell tree takes it 18 s to rebuild) -----------------
    def save(self, file) -> None:
        """Write the tree (mesh, nodes, bb_indices, bb_coords, parameters) to an ``.npz`` file."""
        kind = int(self._tree.info.kind)
        np.savez(
            file, kind=kind, vertices=self.vertices, elements=self._download("elements"), nodes=self.nodes,
            bb_indices=self.bb_indices, bb_coords=self.bb_coords, n_buckets=self.n_buckets, cells_per_leaf=self.cells_per_leaf,
        )  # fmt: skip

    @classmethod
    def load(cls, file):
        """Read a tree written by :meth:`save`: the arrays are uploaded as they are, no build kernels run."""
        with np.load(file) as z:
            kind = int(z["kind"])
            expected = cls._KIND  # a class attribute: subclasses of either tree load like their parent
            if kind != expected:
                raise ValueError(f"{file} holds a {'face' if kind == _lib.CT_KIND_FACES else 'edge'} tree")
            return cls._from_arrays(
                z["vertices"], z["elements"], z["nodes"], z["bb_indices"], z["bb_coords"], int(z["cells_per_leaf"]),
                int(z["n_buckets"]), kind,
            )  # fmt: skip

    @classmethod
    def _from_arrays(cls, vertices, elements, nodes, bb_indices, bb_coords, cells_per_leaf, n_buckets, kind):
        from numba_celltree_b200.cast import cast_bboxes, cast_vertices

        self = cls.__new__(cls)
        vertices = cast_vertices(vertices, copy=True)
        elements = np.ascontiguousarray(elements, dtype=IntDType)
        if elements.ndim != 2:
            raise ValueError("elements must have shape (n_element, n_max_vert)")
        nodes = np.ascontiguousarray(nodes, dtype=NodeDType)
        bb_indices = np.ascontiguousarray(bb_indices, dtype=IntDType)
        bb_coords = cast_bboxes(bb_coords)
        if len(bb_indices) != len(elements) or len(bb_coords) != len(elements):
            raise ValueError("bb_indices and bb_coords must have one row per element")
        handle = ctypes.c_void_p()
        _lib.check(
            _lib.load().ct_tree_from_arrays(
                vertices.ctypes.data, vertices.shape[0], elements.ctypes.data, elements.shape[0], elements.shape[1], kind,
                nodes.ctypes.data, len(nodes), bb_indices.ctypes.data, bb_coords.ctypes.data, int(cells_per_leaf),
                _lib.CT_MEM_HOST, ctypes.byref(handle),
            )
        )  # fmt: skip
        self._tree = DeviceTree(handle.value)
        self.vertices = vertices
        self.n_buckets = n_buckets
        self.cells_per_leaf = cells_per_leaf
        if kind == _lib.CT_KIND_EDGES:
            self.edges = elements
        return self

    # ---- diagnostics (host side, on the mirrors; serial O(n_nodes) stack loops in the reference: query.py:568-664) ----
    # Vectorised per tree level: a tree of 16.7 M cells has 24 of them.
    def _levels(self):
        """Yield the indices of the inner nodes of every level, root first (children follow their parents)."""
        child = self.nodes["child"]
        frontier = np.zeros(1, dtype=np.intp)
        while frontier.size:
            parents = frontier[child[frontier] != -1]
            if parents.size == 0:
                return
            yield parents
            left = child[parents].astype(np.intp)
            frontier = np.concatenate((left, left + 1))

    @property
    def node_bounds(self):
        """Bounds (xmin, xmax, ymin, ymax) of every node: collect_node_bounds, query.py:568-621."""
        nodes = self.nodes
        bounds = np.empty((len(nodes), 4), dtype=FloatDType)
        bounds[0] = self.bbox
        child, dim, Lmax, Rmin = nodes["child"], nodes["dim"].astype(np.intp), nodes["Lmax"], nodes["Rmin"]
        for parents in self._levels():
            left = child[parents].astype(np.intp)
            right = left + 1
            bounds[left] = bounds[parents]
            bounds[right] = bounds[parents]
            bounds[left, 2 * dim[parents] + 1] = Lmax[parents]
            bounds[right, 2 * dim[parents]] = Rmin[parents]
        return bounds

    def validate_node_bounds(self):
        """For every node, whether its children (or its cells' boxes) lie within its bounds: query.py:624-664."""
        nodes = self.nodes
        bounds = self.node_bounds
        child = nodes["child"]
        valid = np.zeros(len(nodes), dtype=bool)

        def contained(a, b):  # box_contained, geometry_utils.py:303-310
            return (a[:, 0] >= b[:, 0]) & (a[:, 1] <= b[:, 1]) & (a[:, 2] >= b[:, 2]) & (a[:, 3] <= b[:, 3])

        reached = np.zeros(len(nodes), dtype=bool)
        reached[0] = True
        for parents in self._levels():
            left = child[parents].astype(np.intp)
            valid[parents] = contained(bounds[left], bounds[parents]) & contained(bounds[left + 1], bounds[parents])
            reached[left] = True
            reached[left + 1] = True
        leaves = np.flatnonzero(reached & (child == -1))
        sizes = nodes["size"][leaves].astype(np.intp)
        valid[leaves] = True  # a leaf without cells is vacuously valid
        if sizes.sum() > 0:
            owner = np.repeat(leaves, sizes)
            first = np.repeat(nodes["ptr"][leaves].astype(np.intp) - (np.cumsum(sizes) - sizes), sizes)
            slots = np.arange(len(owner), dtype=np.intp) + first
            inside = contained(self.bb_coords[self.bb_indices[slots]], bounds[owner])
            valid[leaves] = np.bincount(owner[~inside], minlength=len(nodes))[leaves] == 0
        return valid

    def to_dict_of_lists(self):
        """Children of every node as ``{index: [left, right] or []}``: celltree_base.py:80-114."""
        return {index: [] if left == -1 else [left, left + 1] for index, left in enumerate(self.nodes["child"].tolist())}
