"""profiles/r02_traffic.json entry from an ncu launch list: DRAM bytes per launch of the dominant kernel.
   python tools/ncu_traffic.py <launches.csv> <config> <kernel-substring> <source-label> [existing.json]"""
import csv, io, json, os, sys
fn, config, kern, label = sys.argv[1:5]
out = sys.argv[5] if len(sys.argv) > 5 else None
lines = [l for l in open(fn) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
n, by = 0, 0.0
for r in rows:
    if kern not in r["Kernel Name"]:
        continue
    m = r["Metric Name"]
    v = float(r["Metric Value"].replace(",", ""))
    if m == "gpu__time_duration.sum":
        n += 1
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        by += v * scale[r["Metric Unit"]]
d = {}
if out and os.path.exists(out):
    d = json.load(open(out))
d[config] = {"kernel": kern, "launches": n, "dram_bytes_per_launch": by / max(n, 1), "source": label}
json.dump(d, open(out, "w") if out else sys.stdout, indent=1)
