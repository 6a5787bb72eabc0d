"""
Compile libcelltree_b200.so in-tree with nvcc for sm_100a.

    python -m numba_celltree_b200.build_ext [--force]

Flags that matter for parity: -fmad=false (the reference's Numba/LLVM code has no fused multiply-adds;
bucket membership, on-edge decisions and pair lists depend on the last bit).  fp64 division and sqrt are
IEEE-correct on the GPU by default.  -lineinfo keeps the ncu source page usable.
"""

from __future__ import annotations

import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libcelltree_b200.so"
SOURCES = [CSRC / name for name in ("lib.cu", "points.cu", "boxes.cu", "edges.cu", "build.cu", "algorithms.cu")]
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted((HERE.parent / "include").glob("*.h"))  # every header: a stale .so must never be used

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr",
    "-cudart", "static",
]  # fmt: skip


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, tag: str = "", defines: tuple = ()) -> pathlib.Path:
    """Compile and link the library.  `tag` / `defines` make an experiment variant libcelltree_b200_<tag>.so
    (same sources, extra -D flags) that CELLTREE_B200_LIB can point the binding at."""
    lib = HERE / f"libcelltree_b200_{tag}.so" if tag else LIB
    if not tag and not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = CSRC / (src.stem + (f"_{tag}" if tag else "") + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}")
    link = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", str(lib), *map(str, objs)]
    subprocess.run(link, check=True)
    return lib


if __name__ == "__main__":
    # python -m numba_celltree_b200.build_ext [--force] [-v] [--variant TAG -DNAME=VALUE ...]
    if "--variant" in sys.argv:
        tag = sys.argv[sys.argv.index("--variant") + 1]
        print(build(force=True, verbose="-v" in sys.argv, tag=tag, defines=tuple(a[2:] for a in sys.argv if a.startswith("-D"))))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
