"""Top SASS instructions by stall samples (with the dominant stall reasons and the CUDA line they belong to) of one
kernel of an ncu report:  ncu -i rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > x.csv
   python tools/ncu_sass_top.py x.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; cur = None; f = None
seen = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "": cur = (f, r[0], r[1].strip()[:70]); continue
    addr = r[2]
    if addr in seen: continue          # the same SASS row is listed under every file section it maps to
    seen[addr] = (cur, r)
isam = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
def num(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(num(r[isam]) for _, r in seen.values())
print("total samples", tot, "sass rows", len(seen))
agg = collections.Counter()
for cur, r in seen.values():
    for i, h in st: agg[h] += num(r[i])
print("stall totals:", ", ".join("%s %d" % (h, n) for h, n in agg.most_common(8)))
for addr, (cur, r) in sorted(seen.items(), key=lambda kv: -num(kv[1][1][isam]))[:top]:
    reasons = sorted(((num(r[i]), h[6:]) for i, h in st), reverse=True)[:2]
    print("%6d %5.1f%% ex=%9d  %-42s %-26s %s:%s %s" % (num(r[isam]), 100.0 * num(r[isam]) / max(tot, 1), num(r[iex]), r[3][:42],
          ",".join("%s=%d" % (h, n) for n, h in reasons if n), cur[0], cur[1], cur[2][:40]))
