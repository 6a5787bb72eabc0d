#!/usr/bin/env python
"""
bench.py -- BASELINE.json's headline metric on its headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric: locate_points queries/s.  Workload (config C2 of BASELINE.json / SURVEY.md 8d): 4096 x 4096
structured quad mesh as CellTree2d (16 777 216 cells), 100 000 000 seeded random points per GPU.
One "step" = one locate_points pass over the whole batch.

  value      device-resident throughput: points and results live in HBM, CUDA events around K steps.
  e2e        the same call through the public API with HOST (pinned) buffers: host->device copy of the
             points and device->host copy of the indices inside the timed region.
  roofline   dominant kernel (k_locate_points): algorithmic bytes (918 B/query at C2) x queries
             / its launch duration (CUDA events on the launching stream), against the measured HBM copy peak;
             `traffic` is that kernel's DRAM bytes per launch from the ncu capture named in profiles/traffic.json.
  cpu_baseline  the CPU oracle (C restatement of the reference's algorithm, OpenMP over queries like the
             reference's prange) timed on this box's host cores on a bounded prefix of the same points.

Multi-GPU (torchrun, one rank per GPU): the tree is built on rank 0 and replicated by NCCL broadcast over
NVLink; queries shard by rank with no data-path collective (weak scaling: every rank owns a full batch).

--impl reference times the CPU oracle alone (all host threads) on the same configuration, each step a
bounded prefix of the batch.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "locate_points queries/s"
UNIT = "queries/s"
# 16 B point + 8 B result + 24 node visits x 32 B + 1.5 cells x (4 B index + 16 B face row + 64 B vertices)
ALGORITHMIC_BYTES_PER_QUERY = 918.0
HBM_FALLBACK_GBS = 6650.0


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload():
    nx = env_int("CELLTREE_BENCH_NX", 4096)
    n_points = env_int("CELLTREE_BENCH_POINTS", 100_000_000)
    name = f"C2: {nx}x{nx} structured quad mesh as CellTree2d ({nx * nx} cells), {n_points} random points locate_points"
    return nx, n_points, name


def measured_peak():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    QUERY = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )  # fmt: skip
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


_RESULT_FD = None


def claim_stdout() -> None:
    """Only the result line may go to stdout: libraries that print there (NCCL's version banner) are sent to stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_rate(tree_data, tolerance, points, target_seconds=12.0):
    """Time the CPU oracle's locate_points on a bounded prefix of `points`; returns (q/s, n, seconds, result)."""
    import oracle

    oracle.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1; the baseline uses every core

    probe = min(len(points), 500_000)
    t0 = time.perf_counter()
    oracle.locate_points(points[:probe], tree_data, tolerance)
    dt = time.perf_counter() - t0
    rate = probe / max(dt, 1e-9)
    n = int(min(len(points), max(probe, rate * target_seconds)))
    t0 = time.perf_counter()
    result = oracle.locate_points(points[:n], tree_data, tolerance)
    dt = time.perf_counter() - t0
    return n / dt, n, dt, result


def reference_cpu_baseline(tree_data, tolerance, points, target_seconds=15.0):
    """
    The `cpu_baseline` leg of the GPU arm: the reference's own Numba kernel query.locate_points (query.py:110-117) from
    baseline/_ref on a bounded prefix of this run's points, all host cores.  `tree_data` are the host mirrors of the
    device tree (bit-identical to the arrays the reference builds, tests/test_gpu_fullscale.py), so the 18.5 s serial
    reference build is not repeated here; `--impl reference` does build with the reference.
    Returns (cpu_baseline dict, result array of the prefix) or (None, reason).
    """
    sys.path.insert(0, str(ROOT / "baseline"))
    try:
        import reference as ref_arm

        ref, info = ref_arm.load_reference()
    except Exception as e:  # noqa: BLE001
        return None, f"{type(e).__name__}: {e}"
    from numba_celltree import query
    from numba_celltree.constants import CellTreeData

    data = CellTreeData(*tree_data)
    t0 = time.perf_counter()
    query.locate_points(points[:100_000], data, tolerance)  # JIT compile (or cache load) + thread pool start
    jit_s = time.perf_counter() - t0
    probe = min(len(points), 1_000_000)
    t0 = time.perf_counter()
    query.locate_points(points[:probe], data, tolerance)
    rate = probe / max(time.perf_counter() - t0, 1e-9)
    n = int(min(len(points), max(probe, rate * target_seconds / 3)))
    per_call, best, result = ref_arm.time_calls(lambda: query.locate_points(points[:n], data, tolerance), 3)
    return {
        "value": n / best,
        "unit": UNIT,
        "cores": info["numba_threads"],
        "kind": "reference",
        "sample": f"first {n} of the {len(points)} points, best of 3 calls of query.locate_points ({best:.2f} s), Numba prange",
        "threading_layer": ref_arm.threading_layer(),
        "jit_s": round(jit_s, 1),
        **info,
    }, result


def run_reference(args):
    """
    The reference arm: the UNMODIFIED reference (Numba prange) from baseline/_ref, all host cores, same workload.
    Falls back to the C/OpenMP port of oracle/ only when the reference cannot be imported, and says so.
    """
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from numba_celltree_b200.synthetic import c2_points, quad_mesh

    sys.path.insert(0, str(ROOT / "baseline"))
    import reference as ref_arm

    nx, n_points, name = workload()
    budget_s = float(os.environ.get("CELLTREE_BENCH_REFERENCE_BUDGET_S", 240))
    vertices, faces = quad_mesh(nx, nx)
    info = {}
    fallback = None
    try:
        ref, info = ref_arm.load_reference()
        from numba_celltree import query
        from numba_celltree.celltree_base import default_tolerance

        kind = "reference"
        t0 = time.perf_counter()
        tree = ref.CellTree2d(vertices, faces, -1)
        build_s = time.perf_counter() - t0
        tolerance = default_tolerance(tree.bb_distances[:, 2])
        data = tree.celltree_data

        def kernel_call(pts):
            return query.locate_points(pts, data, tolerance)

        def api_call(pts):
            return tree.locate_points(pts)

    except ImportError as e:
        import oracle

        fallback = f"reference unavailable ({e}); timing the C/OpenMP port of oracle/ instead"
        kind = "port"
        oracle.set_num_threads(host_threads())
        t0 = time.perf_counter()
        tree = oracle.CellTree2d(vertices, faces, -1)
        build_s = time.perf_counter() - t0
        info = {"host_cores": host_threads(), "numba_threads": oracle.num_threads(), "cpu_model": ref_arm.cpu_model()}

        def kernel_call(pts):
            return oracle.locate_points(pts, tree.celltree_data, tree._tolerance)

        api_call = kernel_call

    # JIT compile / thread pool start, then a probe to size the per-step sample so that the run fits its budget
    t0 = time.perf_counter()
    kernel_call(c2_points(100_000))
    jit_s = time.perf_counter() - t0
    probe = c2_points(min(n_points, 2_000_000))
    t0 = time.perf_counter()
    kernel_call(probe)
    rate = len(probe) / max(time.perf_counter() - t0, 1e-9)
    calls = args.steps + args.warmup + 2  # + the API calls below
    forced = os.environ.get("CELLTREE_BENCH_REFERENCE_SAMPLE")
    sample = min(n_points, int(forced)) if forced else max(min(n_points, int(rate * budget_s / calls)), min(n_points, 1_000_000))
    points = c2_points(sample)  # a prefix of the seed-42 stream == the first `sample` points the GPU arm steps over
    for _ in range(args.warmup):
        kernel_call(points)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        result = kernel_call(points)
    dt = time.perf_counter() - t0
    value = args.steps * sample / dt
    _, api_best, api_result = ref_arm.time_calls(lambda: api_call(points), 2)
    whole = sample == n_points
    sample_text = (
        f"all {n_points} seed-42 points per step" if whole else f"first {sample} of the {n_points} seed-42 points per step "
        f"(sized from a {len(probe)}-point probe so that {calls} calls fit {budget_s:.0f} s)"
    )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": name,
            "step": sample_text,
            "timed_call": "numba_celltree.query.locate_points(points, tree.celltree_data, tolerance) (query.py:110-117)"
            if kind == "reference" else "oracle.locate_points (C/OpenMP port)",
            "api_call_queries_per_s": sample / api_best,
            "api_call": "CellTree2d.locate_points(points) (celltree.py:99-128; includes default_tolerance's Python max() over "
            "n_cells per call)",
            "api_equals_kernel_result": bool(np.array_equal(api_result, result)),
            "tree_build_s": round(build_s, 2),
            "jit_s": round(jit_s, 1),
            "found_fraction": float((result >= 0).mean()),
            "result_sha256_16": __import__("hashlib").sha256(np.ascontiguousarray(result).tobytes()).hexdigest()[:16],
        },
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": info.get("numba_threads", host_threads()),
            "kind": kind,
            "sample": sample_text + f", {args.steps} steps",
            "threading_layer": ref_arm.threading_layer() if kind == "reference" else "OpenMP (gcc)",
            **info,
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if fallback:
        line["fallback"] = fallback
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    from numba_celltree_b200 import CellTree2d, _lib
    from numba_celltree_b200 import distributed as ctd
    from numba_celltree_b200.synthetic import quad_mesh

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    lib = _lib.load()
    _lib.check(lib.ct_set_device(local_rank))
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)

    nx, n_points, name = workload()
    args.warmup = max(args.warmup, 3)

    # ---- the tree: built on rank 0 on the device, replicated over NVLink -----------------------------------
    t0 = time.perf_counter()
    if rank == 0:
        vertices, faces = quad_mesh(nx, nx)
        tree = CellTree2d(vertices, faces, -1)
        build_ms_cold = tree.build_ms  # includes first-touch growth of the CUDA memory pool
        del tree
        tree = CellTree2d(vertices, faces, -1)
        build_ms = tree.build_ms
        del faces
    else:
        tree = None
        build_ms = build_ms_cold = None
    if world > 1:
        tree = ctd.broadcast_tree(tree, src=0, device=device)
    setup_s = time.perf_counter() - t0
    tolerance = tree._default_tolerance()

    # ---- this rank's batch of queries (seed 42 + rank), pinned on the host, resident on the device ----------
    rng = np.random.default_rng(42 + rank)
    host_points = torch.empty((n_points, 2), dtype=torch.float64).pin_memory()
    host_np = host_points.numpy()
    rng.random(out=host_np.reshape(-1))  # == default_rng(seed).uniform(0, 1, (n, 2))
    host_out = torch.empty(n_points, dtype=torch.int64).pin_memory()
    dev_points = host_points.to(device, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        import torch.distributed as dist

        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident steps ---------------------------------------------------------------------------------
    # clocks / throttle reasons are sampled from the warm-up on, through the device-resident steps, the per-kernel steps
    # and the end-to-end steps (the device-resident region alone lasts under 0.1 s: too short for more than a sample)
    sampler = ClockSampler(local_rank)
    sampler.start()
    dev_out = None
    for _ in range(args.warmup):
        dev_out = tree.locate_points(dev_points)
    barrier()
    launches0 = lib.ct_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        dev_out = tree.locate_points(dev_points)
    stop.record()
    barrier()
    launches = lib.ct_launch_count() - launches0
    ms_total = max_over_ranks(start.elapsed_time(stop))
    ms_per_step = ms_total / args.steps
    value = world * n_points / (ms_per_step * 1e-3)

    # ---- dominant kernel (the traversal, k_locate_points): CUDA events recorded by the library on the launching
    # stream right around that launch; the Morton ordering (key kernel + radix sort passes) is timed separately.
    import ctypes

    _lib.check(lib.ct_profile_enable(1))
    order_ms, kernel_ms = [], []
    for _ in range(min(args.steps, 5)):
        tree.locate_points(dev_points)
        a, b = ctypes.c_double(), ctypes.c_double()
        _lib.check(lib.ct_profile_last(ctypes.byref(a), ctypes.byref(b)))
        order_ms.append(a.value)
        kernel_ms.append(b.value)
    _lib.check(lib.ct_profile_enable(0))
    kernel_avg_ms = float(np.mean(kernel_ms))
    order_avg_ms = float(np.mean(order_ms))
    peak, peak_source = measured_peak()
    achieved = ALGORITHMIC_BYTES_PER_QUERY * n_points / (kernel_avg_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": "k_locate_points<4,false,9> (entry grid + treelet descent + point-in-polygon; one launch per step)",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_source,
        "algorithmic_bytes_per_query": ALGORITHMIC_BYTES_PER_QUERY,
        "kernel_ms": kernel_avg_ms,
        "kernel_share_of_step": kernel_avg_ms / ms_per_step,
        "morton_order_ms": order_avg_ms,
        "unpermute_ms": ms_per_step - order_avg_ms - kernel_avg_ms,
        "step_achieved": ALGORITHMIC_BYTES_PER_QUERY * n_points / (ms_per_step * 1e-3) / 1e9,
        "step_frac": ALGORITHMIC_BYTES_PER_QUERY * n_points / (ms_per_step * 1e-3) / 1e9 / peak,
    }
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists() and n_points == 100_000_000 and nx == 4096:
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get("k_locate_points_bytes_per_launch")
            roofline["traffic_source"] = "profiles/traffic.json (ncu --set full of this kernel on this workload)"
            roofline["dram_gbs"] = roofline["traffic"] / (kernel_avg_ms * 1e-3) / 1e9
            roofline["dram_frac"] = roofline["dram_gbs"] / peak
        except Exception:
            pass
    roofline["note"] = (
        "frac counts SURVEY 8d's algorithmic bytes (24 node visits x 32 B + cells + point + result per query); most of them "
        "are served by L1/L2 after Morton ordering or skipped by the entry grid, so frac > 1 is not a DRAM rate -- dram_frac "
        "(ncu DRAM bytes / launch time / peak) is; the kernel is issue- and latency-bound"
    )

    # ---- the other half of config C2: locate_points + barycentric (Wachspress) weights, device-resident ---------------------
    weights_line = None
    if rank == 0 and world == 1:
        for _ in range(2):
            tree.compute_barycentric_weights(dev_points)
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        w = None
        for _ in range(3):
            del w  # free the previous result first: the allocator then reuses its block instead of growing
            _, w = tree.compute_barycentric_weights(dev_points)
        w1.record()
        torch.cuda.synchronize()
        w_ms = w0.elapsed_time(w1) / 3
        weights_line = {
            "metric": "compute_barycentric_weights queries/s (device-resident)",
            "value": n_points / (w_ms * 1e-3),
            "ms_per_step": w_ms,
            "weights_row_sum_mean": float(w.sum().item()) / n_points,
        }
        del w

    # ---- end to end through the public API with pinned host buffers --------------------------------------------------
    out_np = host_out.numpy()
    for _ in range(2):
        tree.locate_points(host_np, out=out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tree.locate_points(host_np, out=out_np)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    e2e = {
        "value": world * n_points * args.steps / e2e_s,
        "unit": UNIT,
        "h2d_bytes_per_step": int(n_points * 16),
        "d2h_bytes_per_step": int(n_points * 8),
        "ms_per_step": 1e3 * e2e_s / args.steps,
    }
    same = bool(torch.equal(dev_out.cpu(), host_out))

    # ---- CPU baseline + parity spot check (rank 0, single-GPU run only) ------------------------------------------------
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle

        data = tree.celltree_data  # mirrors of the device tree (bit-identical to the reference's arrays)
        # parity: every result against the oracle (the pinned C restatement; it is the fast checker) ...
        rate, n_cpu, secs, cpu_result = cpu_oracle_rate(data, tolerance, host_np)
        parity = {"checked_queries": n_cpu, "bit_exact": bool(np.array_equal(cpu_result, out_np[:n_cpu])), "checker": "oracle (C port)"}
        port = {
            "value": rate, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"first {n_cpu} of the {n_points} points, {secs:.1f} s, OpenMP static over queries",
        }  # fmt: skip
        # ... and the CPU baseline proper: the reference's own Numba kernel on a bounded prefix of the same points
        cpu_baseline, ref_result = reference_cpu_baseline(data, tolerance, host_np)
        if cpu_baseline is None:
            cpu_baseline = dict(port, fallback=f"reference unavailable: {ref_result}")
        else:
            n_ref = len(ref_result)
            parity["reference_checked_queries"] = n_ref
            parity["reference_bit_exact"] = bool(np.array_equal(ref_result, out_np[:n_ref]))
            cpu_baseline["port_beside_it"] = port
        del cpu_result

    # ---- second headline metric: intersect_faces pairs/s (C5), single GPU -----------------------------------------------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = intersect_faces_metric(torch)

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": name,
                "per_gpu_queries": n_points,
                "sharding": "tree replicated (NCCL broadcast), queries sharded by rank, no data-path collective",
                "l2": "inputs larger than L2 (1.6 GB of points + 1.1 GB of tree per step vs 126 MB)",
                "tree_build_ms": build_ms,
                "tree_build_ms_first_call": build_ms_cold,
                "queries_execution_order": "Morton (Z-order) over the tree bbox, radix sort inside the timed step",
                "setup_s": round(setup_s, 2),
                "tree_depth": tree.depth,
                "tolerance": tolerance,
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "e2e_equals_device_result": same,
            "secondary": secondary,
            "barycentric_weights": weights_line,
        }
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def intersect_faces_metric(torch):
    """C5: 1000 x 1000 quads intersect_faces against the 2M-triangle Delaunay tree (pairs/s, end to end)."""
    from numba_celltree_b200 import CellTree2d
    from numba_celltree_b200.synthetic import delaunay_mesh, quad_mesh

    n_pts = env_int("CELLTREE_BENCH_C3_POINTS", 1_000_000)
    nq = env_int("CELLTREE_BENCH_C5_NX", 1000)
    try:
        vertices, faces = delaunay_mesh(n_pts, seed=1234)
    except Exception as e:  # scipy missing
        return {"unavailable": str(e)}
    tree = CellTree2d(vertices, faces, -1)
    qv, qf = quad_mesh(nq, nq)
    tree.intersect_faces(qv, qf, -1)
    torch.cuda.synchronize()
    times = []
    n_pairs = 0
    area = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        i, j, a = tree.intersect_faces(qv, qf, -1)
        times.append(time.perf_counter() - t0)
        n_pairs = len(i)
        area = float(a.sum())
    best = min(times)
    # the same call with the query mesh already on the device and the pairs left there (what a regridding step that
    # builds its sparse weights on the GPU sees): CUDA events on the current stream, which is the library's
    dqv = torch.from_numpy(qv).cuda()
    dqf = torch.from_numpy(qf).cuda()
    tree.intersect_faces(dqv, dqf, -1)
    torch.cuda.synchronize()
    dev_ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        di, dj, da = tree.intersect_faces(dqv, dqf, -1)
        e1.record()
        torch.cuda.synchronize()
        dev_ms.append(e0.elapsed_time(e1))
    device_resident = {
        "value": int(di.shape[0]) / (min(dev_ms) * 1e-3),
        "unit": "pairs/s",
        "ms": min(dev_ms),
        "pairs_equal_host_path": bool(int(di.shape[0]) == n_pairs and np.array_equal(di.cpu().numpy(), i) and np.array_equal(dj.cpu().numpy(), j)),
    }
    return {
        "metric": "intersect_faces pairs/s",
        "device_resident": device_resident,
        "value": n_pairs / best,
        "unit": "pairs/s",
        "pairs": n_pairs,
        "seconds": best,
        "sum_area": area,
        "workload": f"C5: {nq}x{nq} quads intersect_faces against Delaunay({n_pts} pts) = {len(faces)} triangles; NumPy in, NumPy out",
        "tree_build_ms": tree.build_ms,
        "tree_depth": tree.depth,
    }


if __name__ == "__main__":
    main()
