"""Top stall lines of an ncu source-page CSV: python ncu_src_top.py file.csv [N]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
isrc = hdr.index("Source"); isam = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isam] or 0) for r in data)
print("kernel:", rows[0][1][:100]); print("total samples", tot, "SASS rows", len(data))
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:9])
c = collections.Counter(); s = collections.Counter()
for r in data:
    n = int(r[iex] or 0); c[n] += 1; s[n] += int(r[isam] or 0)
print("exec-count groups (count: #instr, samples):", [(n, k, s[n]) for n, k in sorted(c.items(), key=lambda kv: -s[kv[0]])[:8]])
for r in sorted(data, key=lambda r: -int(r[isam] or 0))[:top_n]:
    st = sorted([(int(r[i] or 0), hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(r[isam].rjust(6), r[iex].rjust(9), r[isrc][:80].ljust(80), st)
