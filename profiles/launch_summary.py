"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections, csv, io, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict(); tot = 0.0; k = 0
for row in csv.DictReader(io.StringIO("".join(lines))):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    k += 1
    if k <= skip: continue
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
    a = agg.setdefault(row["Kernel Name"][:64] + " " + row["Grid Size"], [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:90s} n={c:4d} total={ms:9.3f} ms avg={ms/c:8.4f} ms share={100*ms/tot:5.1f}%")
print("total", round(tot, 3), "ms")
