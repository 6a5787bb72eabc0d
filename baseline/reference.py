"""
The CPU arm of bench.py: the UNMODIFIED reference (Deltares/numba_celltree, Numba `prange`) imported from
baseline/_ref (see vendor_ref.py), timed on this box's host cores.

Only bench.py imports this module.  Nothing of this repository's own code is on the timed path: the tree is the
reference's `CellTree2d`, the timed calls are its array-level kernel `query.locate_points(points, tree.celltree_data,
tolerance)` (query.py:110-117) and its API method `CellTree2d.locate_points` (celltree.py:99-128).
"""

from __future__ import annotations

import os
import pathlib
import sys
import time

HERE = pathlib.Path(__file__).resolve().parent
REF_ROOT = pathlib.Path(os.environ.get("CELLTREE_REFERENCE_ROOT") or HERE / "_ref")  # the override is for the fallback test


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def load_reference():
    """Import numba_celltree from baseline/_ref.  Returns (module, info dict) or raises ImportError with the reason."""
    if not (REF_ROOT / "numba_celltree" / "__init__.py").exists():
        raise ImportError(f"{REF_ROOT}/numba_celltree is missing (run `python baseline/vendor_ref.py` where /root/reference exists)")
    # every host core: torchrun exports OMP_NUM_THREADS=1 to its children, which would also cap Numba's OpenMP layer
    threads = host_threads()
    os.environ.setdefault("NUMBA_NUM_THREADS", str(threads))
    if os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(threads)
    cache = REF_ROOT / ".numba_cache"
    try:
        cache.mkdir(parents=True, exist_ok=True)
        probe = cache / ".write_probe"
        probe.write_text("x")
        probe.unlink()
    except OSError:
        import tempfile

        cache = pathlib.Path(tempfile.mkdtemp(prefix="numba_cache_"))
    os.environ.setdefault("NUMBA_CACHE_DIR", str(cache))
    if str(REF_ROOT) not in sys.path:
        sys.path.insert(0, str(REF_ROOT))
    import numba
    import numba_celltree  # noqa: F401  (the vendored reference)
    from numba_celltree import query  # noqa: F401

    if pathlib.Path(numba_celltree.__file__).resolve().parent != (REF_ROOT / "numba_celltree").resolve():
        raise ImportError(f"numba_celltree was imported from {numba_celltree.__file__}, not from baseline/_ref")
    info = {
        "package": f"numba_celltree {numba_celltree.__version__} (baseline/_ref, unmodified)",
        "numba": numba.__version__,
        "numba_threads": int(numba.get_num_threads()),
        "host_cores": threads,
        "os_cpu_count": os.cpu_count(),
        "cpu_model": cpu_model(),
    }
    return numba_celltree, info


def threading_layer() -> str:
    """Numba's threading layer: only known after the first parallel kernel has run."""
    import numba

    try:
        return str(numba.threading_layer())
    except Exception as e:  # not initialised yet
        return f"unknown ({type(e).__name__})"


def time_calls(fn, repeats: int):
    """Wall time of `repeats` back-to-back calls; returns (seconds per call, best single call, last result)."""
    result = None
    singles = []
    t_all = time.perf_counter()
    for _ in range(repeats):
        t0 = time.perf_counter()
        result = fn()
        singles.append(time.perf_counter() - t0)
    total = time.perf_counter() - t_all
    return total / max(repeats, 1), min(singles) if singles else float("nan"), result
