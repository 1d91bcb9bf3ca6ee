"""Per-CUDA-source-line stall samples of one kernel of an ncu report:
   ncu -i rep --page source --csv --print-source cuda,sass --launch-skip K --launch-count 1 > x.csv; python ncu_lines.py x.csv [N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file = None


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        isam = hdr.index("# Samples"); iex = hdr.index("Instructions Executed")
        continue
    if hdr is None:
        continue
    if r[0] != "":   # CUDA source line row (aggregated)
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [r[1].strip(), 0, 0])
        a[1] += num(r[isam])
        a[2] += num(r[iex])
tot = sum(a[1] for a in agg.values())
print("total samples", tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6d %5.1f%% %10d  %s:%d  %s" % (a[1], 100.0 * a[1] / max(tot, 1), a[2], k[0], k[1], a[0][:110]))
