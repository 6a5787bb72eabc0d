"""
ctypes binding of libcelltree_b200.so (include/celltree_b200.h).

There is no fallback: if the CUDA library is missing or no GPU is usable the calls raise.
"""

from __future__ import annotations

import ctypes
import pathlib

import os

HERE = pathlib.Path(__file__).resolve().parent
# CELLTREE_B200_LIB: an experiment build of the same library (build_ext.build_variant), for A/B measurements
LIB_PATH = pathlib.Path(os.environ.get("CELLTREE_B200_LIB") or HERE / "libcelltree_b200.so")

CT_OK = 0
CT_ERR_CUDA = 1
CT_ERR_VALUE = 2
CT_ERR_UNBUCKETABLE = 3
CT_ERR_DEPTH = 4
CT_ERR_CLIP_STATE = 5
CT_ERR_ZERO_DIVISION = 6
CT_MEM_HOST = 0
CT_MEM_DEVICE = 1
CT_KIND_FACES = 0
CT_KIND_EDGES = 1

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f64 = ctypes.c_double
c_void_p = ctypes.c_void_p


class TreeInfo(ctypes.Structure):
    _fields_ = [
        ("n_vertex", c_i64),
        ("n_elem", c_i64),
        ("n_nodes", c_i64),
        ("n_max_vert", c_i32),
        ("kind", c_i32),
        ("n_buckets", c_i32),
        ("cells_per_leaf", c_i32),
        ("depth", c_i32),
        ("device", c_i32),
        ("bbox", c_f64 * 4),
        ("default_tolerance", c_f64),
        ("build_ms", c_f64),
    ]


_SIGNATURES = {
    "ct_last_error": (ctypes.c_char_p, []),
    "ct_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "ct_set_device": (ctypes.c_int, [ctypes.c_int]),
    "ct_set_stream": (ctypes.c_int, [c_void_p]),
    "ct_launch_count": (c_i64, []),
    "ct_set_sort_bits": (ctypes.c_int, [c_i32]),
    "ct_profile_enable": (ctypes.c_int, [c_i32]),
    "ct_profile_last": (ctypes.c_int, [ctypes.POINTER(c_f64), ctypes.POINTER(c_f64)]),
    "ct_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(c_void_p)]),
    "ct_host_free": (None, [c_void_p]),
    "ct_host_trim": (None, []),
    "ct_device_trim": (None, []),
    "ct_tree_create": (
        ctypes.c_int,
        [c_void_p, c_i64, c_void_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_f64, c_i32, ctypes.POINTER(c_void_p)],
    ),
    "ct_tree_from_arrays": (
        ctypes.c_int,
        [c_void_p, c_i64, c_void_p, c_i64, c_i32, c_i32, c_void_p, c_i64, c_void_p, c_void_p, c_i32, c_i32, ctypes.POINTER(c_void_p)],
    ),
    "ct_tree_get_info": (ctypes.c_int, [c_void_p, ctypes.POINTER(TreeInfo)]),
    "ct_tree_download": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32]),
    "ct_tree_update_nodes": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32]),
    "ct_tree_destroy": (None, [c_void_p]),
    "ct_locate_points": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_f64, c_void_p, c_void_p, c_i32]),
    "ct_locate_boxes": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, c_i32, ctypes.POINTER(c_void_p)]),
    "ct_locate_faces": (
        ctypes.c_int,
        [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i32, c_i64, c_i32, c_i32, c_i32, ctypes.POINTER(c_void_p)],
    ),
    "ct_intersect_edges": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, ctypes.POINTER(c_void_p)]),
    "ct_set_hit_log": (ctypes.c_int, [c_i64]),
    "ct_result_size": (c_i64, [c_void_p]),
    "ct_result_payload_width": (c_i32, [c_void_p]),
    "ct_result_fetch": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32]),
    "ct_result_free": (None, [c_void_p]),
    "ct_liang_barsky_line_box_clip": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_i32]),
    "ct_cohen_sutherland_line_box_clip": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p, c_i32]),
    "ct_cyrus_beck_line_polygon_clip": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i32, c_f64, c_void_p, c_void_p, c_void_p, c_i32]),
    "ct_points_in_polygon": (ctypes.c_int, [c_void_p, c_i64, c_void_p, c_i32, c_void_p, c_i32]),
    "ct_points_in_triangles": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_i32, c_void_p, c_i64, c_f64, c_void_p, c_i32]),
    "ct_profile_binning": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_i32, ctypes.POINTER(c_f64)]),
    "ct_locate_points_stats": (ctypes.c_int, [c_void_p, c_void_p, c_i64, c_f64, ctypes.POINTER(c_i64)]),
    "ct_measure_read_bandwidth": (ctypes.c_int, [ctypes.c_size_t, c_i32, ctypes.POINTER(c_f64)]),
}

EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once). Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m numba_celltree_b200.build_ext` "
                "(nvcc, sm_100a). numba_celltree_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status == CT_OK:
        return
    message = load().ct_last_error().decode("utf-8", "replace")
    if status == CT_ERR_VALUE:
        raise ValueError(message)
    if status == CT_ERR_UNBUCKETABLE:
        raise IndexError(message)
    if status == CT_ERR_ZERO_DIVISION:
        raise ZeroDivisionError(message)
    raise RuntimeError(f"libcelltree_b200 error {status}: {message}")


# ---- result arrays in page-locked host memory -------------------------------------------------------------------
# A device-to-host copy into a fresh pageable ndarray runs at a fraction of the PCIe rate (driver staging plus a
# page fault per 4 KiB); results of at least PINNED_MIN_BYTES are therefore NumPy arrays over page-locked blocks
# that the library recycles (ct_host_alloc / ct_host_free).  CELLTREE_PINNED_RESULTS=0 switches this off.
PINNED_MIN_BYTES = 1 << 20
_pinned_results = os.environ.get("CELLTREE_PINNED_RESULTS", "1") != "0"


def set_pinned_results(enabled: bool) -> None:
    global _pinned_results
    _pinned_results = bool(enabled)


class _PinnedBlock:
    __slots__ = ("address",)

    def __init__(self, address: int):
        self.address = address

    def __del__(self):
        try:
            if _lib is not None and self.address:
                _lib.ct_host_free(self.address)
        except Exception:
            pass


def result_array(shape, dtype):
    """An uninitialised C-contiguous ndarray for a result: page-locked when it is large, plain otherwise."""
    import numpy as np

    dtype = np.dtype(dtype)
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    nbytes = dtype.itemsize
    for x in shape:
        nbytes *= x
    if not _pinned_results or nbytes < PINNED_MIN_BYTES:
        return np.empty(shape, dtype=dtype)
    address = c_void_p()
    if load().ct_host_alloc(nbytes, ctypes.byref(address)) != CT_OK or not address.value:
        return np.empty(shape, dtype=dtype)  # page-locking refused (limits): pageable memory still works
    buffer = (ctypes.c_char * nbytes).from_address(address.value)
    buffer._ct_block = _PinnedBlock(address.value)  # the array's base keeps the block alive
    return np.frombuffer(buffer, dtype=dtype).reshape(shape)
