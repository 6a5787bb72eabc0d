"""profiles/traffic.json from an ncu capture of the point path (profiles/run_ncu.sh): DRAM and L2 bytes per launch of the
four kernels of one locate_points step on C2, stamped with the hash of the kernel sources the capture was taken from
(bench.py drops the figures when the sources that are running hash differently).

    python profiles/make_traffic.py gpurun_out/points_r02.ncu-rep
"""
import csv, io, json, pathlib, subprocess, sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import importlib.util

spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]


def value(row, name):
    i = hdr.index(name)
    v = float(row[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "sector": 32.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
    return v * scale


kernels = {}
for row in rows[2:]:
    name = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
    kernels[name] = {
        "dram_bytes": value(row, "dram__bytes_read.sum") + value(row, "dram__bytes_write.sum"),
        "l2_bytes": value(row, "lts__t_sectors.sum"),
        "duration_ms_under_ncu": value(row, "gpu__time_duration.sum"),
        "local_sectors": (value(row, "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum") + value(row, "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum")) / 32.0,
        "global_sectors": (value(row, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") + value(row, "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum")) / 32.0,
    }
traversal = next(v for k, v in kernels.items() if k.startswith("k_locate_points_binned"))
out = {
    "csrc_sha256": bench.csrc_sha256(),
    "source": f"ncu --set full of one locate_points step on C2 (100 M points, 4096 x 4096 quads): {pathlib.Path(rep).name}, see profiles/run_ncu.sh",
    "traversal_dram_bytes_per_launch": traversal["dram_bytes"],
    "traversal_l2_bytes_per_launch": traversal["l2_bytes"],
    "step_dram_bytes": sum(v["dram_bytes"] for v in kernels.values()),
    "kernels": kernels,
}
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
