"""
The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the unmodified reference under Numba on the host
cores; the C port only as a declared fallback) prints exactly one JSON line on stdout with the agreed keys, on rank 0 only.
Sizes are cut down through the environment so the test takes seconds.
"""

import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def run_bench(*args, **env):
    full_env = dict(os.environ, CELLTREE_BENCH_NX="64", CELLTREE_BENCH_POINTS="200000", CELLTREE_BENCH_REFERENCE_SAMPLE="50000", **env)
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, env=full_env, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    done = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert done.returncode == 0, done.stderr[-2000:]
    lines = [l for l in done.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, done.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "locate_points queries/s" and line["unit"] == "queries/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert "workload" in line["config"] and "64x64" in line["config"]["workload"]
    cpu = line["cpu_baseline"]
    assert cpu["cores"] >= 1 and cpu["value"] == line["value"] and "50000" in cpu["sample"]
    if (ROOT / "baseline" / "_ref" / "numba_celltree").is_dir():
        # the unmodified reference, vendored by baseline/vendor_ref.py, under Numba
        assert cpu["kind"] == "reference" and "fallback" not in line and cpu["numba_threads"] >= 1 and cpu["threading_layer"]
        assert "query.locate_points" in line["config"]["timed_call"] and line["config"]["api_equals_kernel_result"] is True
    else:
        assert cpu["kind"] == "port" and "fallback" in line
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_falls_back_to_the_port_and_says_so():
    done = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", CELLTREE_REFERENCE_ROOT="/nonexistent")
    assert done.returncode == 0, done.stderr[-2000:]
    line = json.loads(done.stdout.strip())
    assert line["cpu_baseline"]["kind"] == "port" and "reference unavailable" in line["fallback"]


def test_reference_arm_is_silent_on_other_ranks():
    done = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", RANK="1", WORLD_SIZE="2")
    assert done.returncode == 0, done.stderr[-2000:]
    assert done.stdout.strip() == ""


def test_traffic_file_is_stamped_with_a_source_hash():
    """profiles/traffic.json carries the hash of the kernel sources its ncu capture was taken from; bench.py uses the figures
    only when the sources that run hash the same (the GPU box has no .git to compare commits with)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    stamp = json.loads((ROOT / "profiles" / "traffic.json").read_text())
    assert len(stamp["csrc_sha256"]) == 16 and len(bench.csrc_sha256()) == 16
    assert stamp["traversal_dram_bytes_per_launch"] > 0 and stamp["step_dram_bytes"] >= stamp["traversal_dram_bytes_per_launch"]
    assert any(name.startswith("k_locate_points_binned") for name in stamp["kernels"])
