#!/usr/bin/env python
"""
Vendor the UNMODIFIED reference package for the CPU arm of bench.py.

    python baseline/vendor_ref.py

Copies /root/reference/numba_celltree (pure Python + Numba, nothing to compile) to baseline/_ref/numba_celltree.
baseline/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so it travels to the
GPU box, where /root/reference does not exist.  `bench.py --impl reference` and the `cpu_baseline` leg import it from
there (`kind: "reference"`); nothing in numba_celltree_b200/ does.
"""

from __future__ import annotations

import pathlib
import shutil
import sys

HERE = pathlib.Path(__file__).resolve().parent
SOURCE = pathlib.Path("/root/reference/numba_celltree")
TARGET = HERE / "_ref" / "numba_celltree"


def vendor(force: bool = False) -> bool:
    """Returns True when baseline/_ref/numba_celltree exists afterwards."""
    if not SOURCE.is_dir():
        return TARGET.is_dir()
    if TARGET.is_dir() and not force:
        newest_src = max(p.stat().st_mtime for p in SOURCE.rglob("*.py"))
        newest_dst = max((p.stat().st_mtime for p in TARGET.rglob("*.py")), default=0.0)
        if newest_dst >= newest_src:
            return True
    if TARGET.exists():
        shutil.rmtree(TARGET)
    TARGET.parent.mkdir(parents=True, exist_ok=True)
    shutil.copytree(SOURCE, TARGET, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.nbi", "*.nbc"))
    return True


if __name__ == "__main__":
    ok = vendor(force="--force" in sys.argv)
    print(f"{TARGET}: {'ready' if ok else 'unavailable (no /root/reference here and no earlier copy)'}")
