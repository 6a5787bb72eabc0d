"""No GPU needed: the C-ABI library loads, exports every symbol include/celltree_b200.h declares, and fails loudly."""

import ctypes
import pathlib
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "celltree_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ct_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from numba_celltree_b200 import _lib, build_ext

    build_ext.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/celltree_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), "ctypes signatures and header disagree"


def test_node41_layout_matches_numpy_dtype():
    from numba_celltree_b200.constants import NodeDType

    assert NodeDType.itemsize == 41
    assert [NodeDType.fields[k][1] for k in ("child", "Lmax", "Rmin", "ptr", "size", "dim")] == [0, 8, 16, 24, 32, 40]


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from numba_celltree_b200 import CellTree2d

    with pytest.raises(RuntimeError, match="CUDA"):
        CellTree2d(np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0, 1, 2]]), -1)


def test_product_package_does_not_import_the_oracle():
    for path in (ROOT / "numba_celltree_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    for path in (ROOT / "numba_celltree_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(errors="ignore"), path


def test_argument_validation_happens_before_the_device_is_touched():
    from numba_celltree_b200 import CellTree2d, EdgeCellTree2d

    v = [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]
    with pytest.raises(ValueError):
        CellTree2d(v, [[0, 1, 2]], -1, n_buckets=1)
    with pytest.raises(ValueError):
        CellTree2d(v, [[0, 1, 2]], -1, cells_per_leaf=0)
    with pytest.raises(ValueError):
        CellTree2d([[0.0, 0.0, 0.0]], [[0, 1, 2]], -1)
    with pytest.raises(ValueError):
        CellTree2d(v, [0, 1, 2], -1)
    with pytest.raises(ValueError):
        CellTree2d(v, np.zeros((1, 33), dtype=int), -1)
    with pytest.raises(ValueError):
        EdgeCellTree2d(v, [[0, 1]], n_buckets=1)


def test_result_arrays_fall_back_to_pageable_memory_when_pinning_is_refused():
    """Without a CUDA device page-locking fails: results are then plain ndarrays (memory only -- compute has no fallback)."""
    import torch

    from numba_celltree_b200 import _lib

    a = _lib.result_array((1 << 18,), np.int64)  # 2 MiB: above the pinning threshold
    assert a.shape == (1 << 18,) and a.dtype == np.int64 and a.flags.writeable and a.flags.c_contiguous
    a[:] = 7
    b = _lib.result_array((3, 5), np.float64)  # small: always plain
    assert b.flags.owndata and b.shape == (3, 5)
    if not torch.cuda.is_available():
        assert a.flags.owndata
