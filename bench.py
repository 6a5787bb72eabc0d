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
  roofline   dominant kernel (the traversal, k_locate_points_binned): the bytes that kernel as built reads and writes per
             query -- counted by an instrumented run of the same walk (ct_locate_points_stats) -- x queries / its launch
             duration (CUDA events on the launching stream), against the measured HBM peak; the L2 and HBM read rates are
             measured in the run as well; `traffic` is the kernel's DRAM bytes per launch from the ncu capture named in
             profiles/traffic.json, used only if that capture was taken from the kernel sources that are running.
  cpu_baseline  the reference itself (numba_celltree under Numba, vendored to baseline/_ref by baseline/vendor_ref.py): its
             query.locate_points on a bounded prefix of the same points, all host cores; the C port of oracle/ beside it.
  secondary  driver-run lines for the other BASELINE.json configurations (C2 weights, C3 boxes, C4 segments, C5 faces, C1 on
             the CPU reference).

Multi-GPU (torchrun, one rank per GPU): the tree is built on rank 0 and replicated by NCCL broadcast over NVLink.  `value`
is weak scaling (every rank owns a full batch, no data-path collective); `strong_scaling` splits the ONE batch by
shard_range and assembles the results on every GPU; `multi_gpu` times the sharded variable-length calls with their one
all-gather; `parity_multi` compares rank 0 alone with the assembled shards.

--impl reference times the UNMODIFIED reference (Numba prange, all host cores) on the same configuration; the C/OpenMP
port only if the reference cannot be imported, and then says so.
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


def csrc_sha256() -> str:
    """Hash of the kernel sources: ties a committed ncu capture (profiles/traffic.json) to the code that is running
    (the GPU box has no .git, so a commit hash cannot be compared there)."""
    import hashlib

    h = hashlib.sha256()
    files = sorted((ROOT / "numba_celltree_b200" / "csrc").glob("*.cu*")) + sorted((ROOT / "include").glob("*.h"))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def device_ms(torch, fn, reps=3):
    """Best device time of `fn` over `reps` runs (CUDA events on the current stream, which is the library's)."""
    fn()
    torch.cuda.synchronize()
    best, result = None, None
    for _ in range(reps):
        result = None  # free the previous result first: the allocator then reuses its block
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        result = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, result


def host_ms(torch, fn, reps=2):
    """Best wall time of `fn` (NumPy in, NumPy out: the copies are inside)."""
    fn()
    best, result = None, None
    for _ in range(reps):
        result = None
        t0 = time.perf_counter()
        result = fn()
        torch.cuda.synchronize()
        ms = 1e3 * (time.perf_counter() - t0)
        best = ms if best is None else min(best, ms)
    return best, result


def traversal_roofline(lib, _lib, tree, dev_points, tolerance, n_points, nx, kernel_ms, order_ms, ms_per_step, peak, peak_source):
    """The roofline object of the dominant kernel, from what that kernel as built has to move."""
    import ctypes

    m = int(tree._tree.info.n_max_vert)
    sample = min(n_points, 4_000_000)
    stats = (ctypes.c_int64 * 6)()
    _lib.check(lib.ct_locate_points_stats(tree._tree.handle, dev_points.data_ptr(), sample, float(tolerance), stats))
    slots, headers, cells, pushes, from_grid, found = (stats[k] / sample for k in range(6))
    # per query: its 32-byte record, the entry-grid lookup (4-byte handle + the two 8-byte cell bounds), 16 bytes per node slot
    # and per treelet header read, one row of elem_xy (16 bytes per vertex) per cell tested, the 8-byte (index, result) pair
    bytes_per_query = 32.0 + 4.0 + 16.0 + 16.0 * slots + 16.0 * headers + cells * 16.0 * m + 8.0
    achieved = bytes_per_query * n_points / (kernel_ms * 1e-3) / 1e9
    info = tree._tree.info
    tree_bytes = 128.0 * (int(info.n_nodes) / 3.5) + int(info.n_elem) * 16.0 * m  # treelets (about 3.5 nodes per line) + elem_xy
    compulsory = 16.0 * n_points + 8.0 * n_points + tree_bytes  # points in, results out, tree once
    l2 = ctypes.c_double()
    hbm = ctypes.c_double()
    _lib.check(lib.ct_measure_read_bandwidth(64 << 20, 40, ctypes.byref(l2)))
    _lib.check(lib.ct_measure_read_bandwidth(4 << 30, 2, ctypes.byref(hbm)))
    roofline = {
        "bound": "hbm",
        "kernel": "k_locate_points_binned<4,false,4,false> (tile sort + entry grid + treelet descent + point-in-polygon; one launch per step)",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_source,
        "algorithmic_bytes_per_query": bytes_per_query,
        "algorithmic_bytes_formula": "32 record + 4 entry handle + 16 entry bounds + 16*slots + 16*headers + cells*16*M + 8 result pair",
        "per_query": {"node_slots": slots, "treelet_headers": headers, "cells_tested": cells, "stack_pushes": pushes,
                      "started_from_entry_grid": from_grid, "found": found, "sample": sample},
        "kernel_ms": kernel_ms,
        "kernel_share_of_step": kernel_ms / ms_per_step,
        "binning_ms": order_ms,
        "results_to_caller_order_ms": ms_per_step - order_ms - kernel_ms,
        "step_compulsory_hbm_bytes": compulsory,
        "step_compulsory_frac": compulsory / (ms_per_step * 1e-3) / 1e9 / peak,
        "measured_in_run": {"l2_read_gbs": l2.value, "hbm_read_gbs": hbm.value,
                            "how": "ct_measure_read_bandwidth: 16-byte loads, 64 MB buffer x 40 sweeps (L2) and 4 GB x 2 sweeps (HBM)"},
        "kernel_frac_of_l2": achieved / l2.value,
    }  # fmt: skip
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists() and n_points == 100_000_000 and nx == 4096:
        try:
            t = json.loads(traffic_file.read_text())
            if t.get("csrc_sha256") == csrc_sha256():
                roofline["traffic"] = t["traversal_dram_bytes_per_launch"]
                roofline["traffic_source"] = t.get("source")
                roofline["dram_gbs"] = roofline["traffic"] / (kernel_ms * 1e-3) / 1e9
                roofline["dram_frac"] = roofline["dram_gbs"] / peak
                if t.get("traversal_l2_bytes_per_launch"):
                    roofline["l2_gbs"] = t["traversal_l2_bytes_per_launch"] / (kernel_ms * 1e-3) / 1e9
                    roofline["l2_frac"] = roofline["l2_gbs"] / l2.value
                if t.get("step_dram_bytes"):
                    roofline["step_dram_bytes"] = t["step_dram_bytes"]
                    roofline["step_dram_over_compulsory"] = t["step_dram_bytes"] / compulsory
            else:
                roofline["traffic_note"] = "profiles/traffic.json was captured from other kernel sources (csrc_sha256 differs): dropped"
        except Exception as e:  # noqa: BLE001
            roofline["traffic_note"] = f"profiles/traffic.json unreadable: {e}"
    roofline["note"] = (
        "frac = bytes the traversal as built reads and writes per query (counted by an instrumented run of the same walk, "
        "ct_locate_points_stats) x queries / launch time / HBM peak; after binning most node and polygon reads are L1/L2 hits, "
        "so kernel_frac_of_l2 and dram_frac (ncu) say where the launch stands against either level"
    )
    return roofline


def secondary_metrics(torch, tree_c2, dev_points, n_points, peak):
    """Driver-run lines for the other BASELINE.json configurations (single GPU): device-resident and NumPy-to-NumPy times,
    the figure of SURVEY 8(d)'s bytes-per-unit formula against the HBM peak."""
    from numba_celltree_b200 import CellTree2d
    from numba_celltree_b200.synthetic import c3_boxes, c4_edges, delaunay_mesh, quad_mesh

    out = []

    def line(config, call, units, unit, dev_ms, e2e_ms, bytes_per_unit, formula, extra=None):
        entry = {
            "config": config, "call": call, "units": units, "unit": unit,
            "device_ms": dev_ms, "device_units_per_s": units / (dev_ms * 1e-3),
            "e2e_ms": e2e_ms, "e2e_units_per_s": None if e2e_ms is None else units / (e2e_ms * 1e-3),
            "roofline": {"bytes_per_unit": bytes_per_unit, "formula": formula, "achieved_gbs": bytes_per_unit * units / (dev_ms * 1e-3) / 1e9,
                         "frac_of_hbm_peak": bytes_per_unit * units / (dev_ms * 1e-3) / 1e9 / peak},
        }  # fmt: skip
        entry.update(extra or {})
        out.append(entry)

    # C2, second half: locate_points + Wachspress weights, device-resident
    w_ms, w = device_ms(torch, lambda: tree_c2.compute_barycentric_weights(dev_points))
    line("C2", "compute_barycentric_weights", n_points, "queries", w_ms, None, 32 + 4 + 16 + 16 * 3 + 1.5 * 64 + 8 + 32,
         "the locate_points bytes + 8*M bytes of weights written per query", {"weights_row_sum_mean": float(w[1].sum().item()) / n_points})
    del w

    n_c3 = env_int("CELLTREE_BENCH_C3_POINTS", 1_000_000)
    n_q = env_int("CELLTREE_BENCH_C3_QUERIES", 10_000_000)
    nq5 = env_int("CELLTREE_BENCH_C5_NX", 1000)
    try:
        vertices, faces = delaunay_mesh(n_c3, seed=1234)
    except Exception as e:  # scipy missing
        return [{"unavailable": str(e)}]
    tree = CellTree2d(vertices, faces, -1)
    mesh = f"Delaunay({n_c3} points) = {len(faces)} triangles, depth {tree.depth}, build {tree.build_ms:.1f} ms"
    boxes = c3_boxes(len(faces), n_q)
    d_boxes = torch.from_numpy(boxes).cuda()
    d_ms, r = device_ms(torch, lambda: tree.locate_boxes(d_boxes))
    pairs = int(r[0].shape[0])
    del r
    h_ms, _ = host_ms(torch, lambda: tree.locate_boxes(boxes))
    line("C3", "locate_boxes", n_q, "boxes", d_ms, h_ms, 7800.0, "2 passes x (32 + 32*V + 36*C) + 16*P with V=87.0, C=26.9, P=12.64 (SURVEY 8d)",
         {"pairs": pairs, "device_pairs_per_s": pairs / (d_ms * 1e-3), "mesh": mesh})
    d_ms, r = device_ms(torch, lambda: tree.intersect_boxes(d_boxes))
    kept = int(r[0].shape[0])
    del r
    h_ms, _ = host_ms(torch, lambda: tree.intersect_boxes(boxes))
    line("C3", "intersect_boxes", n_q, "boxes", d_ms, h_ms, 7800.0 + 116.0 * pairs / n_q, "locate_boxes + 116 B per shortlisted pair for the clip (SURVEY 8d)",
         {"pairs": kept, "device_pairs_per_s": kept / (d_ms * 1e-3)})
    del d_boxes, boxes
    edges = c4_edges(len(faces), n_q)
    d_edges = torch.from_numpy(edges).cuda()
    d_ms, r = device_ms(torch, lambda: tree.intersect_edges(d_edges))
    pairs = int(r[0].shape[0])
    del r
    h_ms, _ = host_ms(torch, lambda: tree.intersect_edges(edges))
    line("C4", "intersect_edges", n_q, "segments", d_ms, h_ms, 11500.0, "2 x (32 + 32*V + 36*C + 60*C') + 48*P + sort, V=109.5, C=41.0, P=8.65 (SURVEY 8d)",
         {"pairs": pairs, "device_pairs_per_s": pairs / (d_ms * 1e-3)})
    del d_edges, edges
    qv, qf = quad_mesh(nq5, nq5)
    dqv, dqf = torch.from_numpy(qv).cuda(), torch.from_numpy(qf).cuda()
    d_ms, r = device_ms(torch, lambda: tree.intersect_faces(dqv, dqf, -1))
    di, dj = r[0].cpu().numpy(), r[1].cpu().numpy()
    del r
    h_ms, (i, j, a) = host_ms(torch, lambda: tree.intersect_faces(qv, qf, -1), reps=3)
    pairs = len(i)
    line("C5", "intersect_faces", pairs, "pairs", d_ms, h_ms, 7800.0 * len(qf) / max(pairs, 1) + 157.0 * 1.21 + 164.0,
         "per final pair: the shortlist walk of its face (7.8 KB per face) + 157 B per SAT pair + 164 B per clipped pair (SURVEY 8d)",
         {"metric": "intersect_faces pairs/s", "query_faces": len(qf), "sum_area": float(a.sum()),
          "pairs_equal_host_path": bool(np.array_equal(di, i) and np.array_equal(dj, j)),
          "workload": f"C5: {nq5}x{nq5} quads intersect_faces against {mesh}"})
    return out


def c1_reference_line(torch=None):
    """C1 of BASELINE.json: the reference itself on its own CPU-runnable case (generate_disk(25, 20) + 1 M points), and this
    library on the same case beside it (NumPy to NumPy, and device-resident), answers compared."""
    sys.path.insert(0, str(ROOT / "baseline"))
    try:
        import reference as ref_arm

        ref, info = ref_arm.load_reference()
    except Exception as e:  # noqa: BLE001
        return {"config": "C1", "unavailable": f"{type(e).__name__}: {e}"}
    from numba_celltree_b200.synthetic import c1_points, generate_disk

    v, f = generate_disk(25, 20)
    tree = ref.CellTree2d(v, f, -1)
    pts = c1_points()
    tree.locate_points(pts[:1000])
    _, best, result = ref_arm.time_calls(lambda: tree.locate_points(pts), 3)
    line = {
        "config": "C1", "call": "numba_celltree.CellTree2d.locate_points (reference, CPU)", "units": len(pts), "unit": "queries",
        "queries_per_s": len(pts) / best, "cores": info["numba_threads"], "cells": len(f), "found_fraction": float((result >= 0).mean()),
    }  # fmt: skip
    if torch is not None:
        import numpy as np

        from numba_celltree_b200 import CellTree2d

        ours = CellTree2d(v, f, -1)
        e2e, found = host_ms(torch, lambda: ours.locate_points(pts), reps=5)
        dev_pts = torch.from_numpy(pts).cuda()
        dev, _ = device_ms(torch, lambda: ours.locate_points(dev_pts), reps=5)
        line["this_library"] = {
            "e2e_ms": e2e, "e2e_queries_per_s": len(pts) / (e2e * 1e-3), "device_ms": dev, "device_queries_per_s": len(pts) / (dev * 1e-3),
            "same_answers_as_the_reference": bool(np.array_equal(found, result)),
        }  # fmt: skip
    return line


def multi_gpu_section(torch, dist, ctd, tree, device, rank, world, n_points, steps, tolerance):
    """world > 1: strong scaling of the one C2 batch, the variable-length sharded calls on C3 / C5 with their one
    collective, and a rank-0-versus-shards parity check."""
    from numba_celltree_b200 import CellTree2d
    from numba_celltree_b200.synthetic import c2_points, c3_boxes, delaunay_mesh, quad_mesh

    def synced_ms(fn, reps):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        result = None
        for _ in range(reps):
            result = None
            result = fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), result

    # ---- strong scaling: the ONE batch of n_points seed-42 points, split by shard_range, results assembled on every GPU
    lo, hi = ctd.shard_range(n_points, rank, world)
    whole = c2_points(n_points)
    mine = torch.from_numpy(whole[lo:hi]).to(device)
    del whole
    equal_shards = n_points % world == 0
    assembled = torch.empty(n_points, dtype=torch.int64, device=device) if equal_shards else None

    def strong_step():
        found = tree.locate_points(mine)
        if assembled is not None:
            ctd.assemble_indices(found, out=assembled)
        return found

    ms, found = synced_ms(strong_step, steps)
    checksum = int(assembled.sum().item()) if assembled is not None else None
    sharded_ms, _ = synced_ms(lambda: tree.locate_points(mine), steps)  # the same step without the assembly
    strong = {
        "metric": METRIC, "scaling": "strong", "total_queries": n_points, "per_gpu_queries": hi - lo, "ms_per_step": ms,
        "value": n_points / (ms * 1e-3), "unit": UNIT,
        "assembled": "distributed.assemble_indices: results narrowed to int32, one all_gather_into_tensor (NCCL), widened to int64 on every GPU, inside the timed step" if equal_shards else "left sharded",
        "ms_per_step_left_sharded": sharded_ms,
        "assembly_bytes_received_per_gpu": (n_points - (hi - lo)) * 4 if equal_shards else 0,
        "result_checksum": checksum,
    }  # fmt: skip
    del mine, assembled, found

    # ---- variable-length calls, sharded: C3 locate_boxes and C5 intersect_faces, exchange_totals inside the timed call
    n_c3 = env_int("CELLTREE_BENCH_C3_POINTS", 1_000_000)
    n_q = env_int("CELLTREE_BENCH_C3_QUERIES", 10_000_000)
    nq5 = env_int("CELLTREE_BENCH_C5_NX", 1000)
    c3 = None
    n_tri = [0]
    if rank == 0:
        vertices, faces = delaunay_mesh(n_c3, seed=1234)
        c3 = CellTree2d(vertices, faces, -1)
        n_tri = [len(faces)]
    dist.broadcast_object_list(n_tri, src=0)
    c3 = ctd.broadcast_tree(c3, src=0, device=device)
    boxes = c3_boxes(n_tri[0], n_q)
    b_lo, b_hi = ctd.shard_range(n_q, rank, world)
    d_boxes = torch.from_numpy(boxes).to(device)  # every rank holds the query set; it works on rows [lo, hi)
    ms_boxes, pieces = synced_ms(lambda: ctd.query_pairs_sharded(c3, "locate_boxes", d_boxes, device=device), 3)
    total_pairs = pieces[4]
    qv, qf = quad_mesh(nq5, nq5)
    dqv, dqf = torch.from_numpy(qv).to(device), torch.from_numpy(qf).to(device)
    ms_faces, fpieces = synced_ms(lambda: ctd.intersect_faces_sharded(c3, dqv, dqf, -1, device=device), 3)
    multi = {
        "locate_boxes_sharded": {"config": "C3", "boxes": n_q, "pairs": int(total_pairs), "ms": ms_boxes, "boxes_per_s": n_q / (ms_boxes * 1e-3),
                                 "pairs_per_s": total_pairs / (ms_boxes * 1e-3), "collective": "one all_gather_into_tensor of the per-rank pair counts"},
        "intersect_faces_sharded": {"config": "C5", "faces": len(qf), "pairs": int(fpieces[4]), "ms": ms_faces,
                                    "pairs_per_s": fpieces[4] / (ms_faces * 1e-3)},
    }  # fmt: skip

    # ---- parity: rank 0 alone versus the assembled shards (what tests/test_gpu_multi.py checks, run by the driver here)
    n_par = 1_000_000
    par_pts = torch.from_numpy(c2_points(n_par)).to(device)
    p_lo, p_hi, part = ctd.locate_points_sharded(tree, par_pts)
    equal = n_par % world == 0
    gathered = torch.empty(n_par, dtype=torch.int64, device=device)
    if equal:
        dist.all_gather_into_tensor(gathered, part)
    n_pb = 200_000
    pb = ctd.query_pairs_sharded(c3, "intersect_boxes", d_boxes[:n_pb], device=device)
    got = ctd.gather_pairs(*pb, dst=0)
    fgot = ctd.gather_pairs(*fpieces, dst=0)
    parity_multi = None
    if rank == 0:
        alone = tree.locate_points(par_pts)
        wi, wj, wa = c3.intersect_boxes(d_boxes[:n_pb])
        fi, fj, fa = c3.intersect_faces(dqv, dqf, -1)
        parity_multi = {
            "locate_points": {"queries": n_par, "bit_exact": bool(equal and torch.equal(alone, gathered))},
            "intersect_boxes": {"boxes": n_pb, "pairs": int(wi.shape[0]),
                                "bit_exact": bool(torch.equal(got[0], wi) and torch.equal(got[1], wj) and torch.equal(got[2], wa))},
            "intersect_faces": {"pairs": int(fi.shape[0]),
                                "bit_exact": bool(torch.equal(fgot[0], fi) and torch.equal(fgot[1], fj) and torch.equal(fgot[2], fa))},
            "how": "rank 0 answers alone; the shards of all ranks are assembled on rank 0 (all_gather / point-to-point at the exchanged offsets)",
        }  # fmt: skip
    return strong, multi, parity_multi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-multi", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import ctypes

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
    dist = None
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
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
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

    # ---- dominant kernel (the traversal): CUDA events recorded by the library on the launching stream right around that
    # launch; the binning before it is timed the same way, the rest of the step is the return of the results
    _lib.check(lib.ct_profile_enable(1))
    order_ms, kernel_ms = [], []
    for _ in range(min(args.steps, 5)):
        tree.locate_points(dev_points)
        a, b = ctypes.c_double(), ctypes.c_double()
        _lib.check(lib.ct_profile_last(ctypes.byref(a), ctypes.byref(b)))
        order_ms.append(a.value)
        kernel_ms.append(b.value)
    _lib.check(lib.ct_profile_enable(0))
    peak, peak_source = measured_peak()
    roofline = traversal_roofline(
        lib, _lib, tree, dev_points, tolerance, n_points, nx, float(np.mean(kernel_ms)), float(np.mean(order_ms)), ms_per_step, peak, peak_source
    )

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
    # the floor of that figure: the two copies alone (pinned host <-> device, both directions at once, every rank at the
    # same time) -- what the host side of this box can move, whatever the kernels do
    copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()

    def bare_copies():
        with torch.cuda.stream(copy_in):
            dev_points.copy_(host_points, non_blocking=True)
        with torch.cuda.stream(copy_out):
            host_out.copy_(dev_out, non_blocking=True)
        torch.cuda.synchronize()

    bare_copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        bare_copies()
    e2e["host_copy_floor_ms"] = 1e3 * max_over_ranks(time.perf_counter() - t0) / 3
    e2e["overhead_over_copies"] = e2e["ms_per_step"] / e2e["host_copy_floor_ms"] - 1.0

    # ---- CPU baseline + parity (rank 0, single-GPU run only) ------------------------------------------------
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
        tree.__dict__.pop("_mirrors", None)  # the 0.7 GB node mirror is not needed any more (nor its checksum per call)

    # ---- the other BASELINE.json configurations, single GPU --------------------------------------------------------
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = secondary_metrics(torch, tree, dev_points, n_points, peak)
        if not args.no_cpu_baseline:
            secondary.append(c1_reference_line(torch))

    # ---- multi-GPU: strong scaling, sharded variable-length calls, parity --------------------------------------------
    strong = multi = parity_multi = None
    if world > 1 and not args.no_multi:
        del dev_points, dev_out
        strong, multi, parity_multi = multi_gpu_section(torch, dist, ctd, tree, device, rank, world, n_points, min(args.steps, 5), tolerance)

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
                "l2": "inputs larger than L2 (1.6 GB of points + 1.4 GB of tree per step vs 126 MB)",
                "tree_build_ms": build_ms,
                "tree_build_ms_first_call": build_ms_cold,
                "queries_execution_order": "every point appended to the slab of its Z-order bin as a 32-byte record (one scatter pass), one "
                "tile per bin sorted in shared memory; results returned through per-window queues; all inside the timed step",
                "setup_s": round(setup_s, 2),
                "tree_depth": tree.depth,
                "tolerance": tolerance,
                "csrc_sha256": csrc_sha256(),
            },
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "e2e_equals_device_result": same,
            "secondary": secondary,
            "strong_scaling": strong,
            "multi_gpu": multi,
            "parity_multi": parity_multi,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
