"""Per-source-line instruction and stall-sample shares of the kernels in an .ncu-rep (read offline with `ncu -i`).

    python profiles/ncu_lines.py gpurun_out/<name>.ncu-rep [top]

Needs a capture taken with `--set full --import-source on` of code built with -lineinfo.  For every source line:
share of the warp instructions executed, average active lanes per instruction, share of the stall samples."""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(text)))
current, header, lines = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        current = r[1].split("/")[-1]
    elif r[0] == "Line No":
        header = r
        i_inst, i_thread, i_samples = header.index("Instructions Executed"), header.index("Thread Instructions Executed"), header.index("# Samples")
    elif header is not None and r[0].isdigit() and len(r) >= len(header) - 5:
        try:
            lines[(current, int(r[0]))] = (int(r[i_inst]), int(r[i_thread]), int(r[i_samples]), r[1].strip())
        except ValueError:
            pass
total = sum(v[0] for v in lines.values()) or 1
samples = sum(v[2] for v in lines.values()) or 1
print(f"{total} warp instructions, {samples} stall samples")
files = {}
for (f, _), v in lines.items():
    a = files.setdefault(f, [0, 0, 0])
    a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
for f, a in sorted(files.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} instructions {100 * a[0] / total:5.1f} %  lanes {a[1] / max(a[0], 1):5.1f}  samples {100 * a[2] / samples:5.1f} %")
for title, key in (("by instructions", 0), ("by stall samples", 2)):
    print(f"--- top {top} lines {title}")
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][key])[:top]:
        print(f"{f}:{l:<4d} inst {100 * v[0] / total:5.2f} %  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / samples:5.2f} %  {v[3][:110]}")
