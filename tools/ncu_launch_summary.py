"""Per-kernel summary (launches, device time, share, DRAM bytes) of an ncu launch list taken with
   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
   python tools/ncu_launch_summary.py profiles/xxx.csv > profiles/xxx_summary.md"""
import collections
import csv
import io
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = collections.OrderedDict()
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void ", "")[:64]
    a = agg.setdefault(k, {"n": 0, "t": 0.0, "rd": 0.0, "wr": 0.0})
    m, u = r["Metric Name"], r["Metric Unit"]
    v = float(r["Metric Value"].replace(",", ""))
    if m == "gpu__time_duration.sum":
        a["n"] += 1
        a["t"] += v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    elif m == "dram__bytes_read.sum":
        a["rd"] += v * scale[u]
    elif m == "dram__bytes_write.sum":
        a["wr"] += v * scale[u]
tot = sum(a["t"] for a in agg.values())
print("| kernel | launches | total ms | share | DRAM read GB | DRAM write GB | DRAM MB / launch |")
print("|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1]["t"]):
    print("| %s | %d | %.3f | %.3f | %.2f | %.2f | %.1f |" % (k, a["n"], a["t"], a["t"] / tot, a["rd"] / 1e9, a["wr"] / 1e9,
                                                          (a["rd"] + a["wr"]) / 1e6 / max(a["n"], 1)))
print("\nTotal %.1f ms over %d launches (cold-cache, serialised: compare shares, not absolutes)." % (tot, sum(a["n"] for a in agg.values())))

if "--traffic-json" in sys.argv:
    import json
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    g = [a for k, a in agg.items() if "gemm_tcgen05_kernel" in k]
    n = sum(a["n"] for a in g)
    json.dump({"kernel": "gemm_tcgen05_kernel (all instantiations)", "launches": n,
               "dram_bytes_per_launch": (sum(a["rd"] + a["wr"] for a in g) / max(n, 1)),
               "dram_read_gb_per_step": sum(a["rd"] for a in g) / 1e9, "dram_write_gb_per_step": sum(a["wr"] for a in g) / 1e9,
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, " + sys.argv[1]}, open(out, "w"))
