"""Aggregate a bench.py --profile-csv per-launch table by (phase, label)."""
import collections, csv, sys
rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
for r in rows:
    k = (r['phase'], r['kind'], r['label'])
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(r['ms']); a[2] += float(r['gflop'])
tot = sum(a[1] for a in agg.values())
print("total ms", round(tot, 2), "launches", len(rows))
for (ph, kind, lab), (n, ms, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{ph:5s} k{kind} n={n:3d} ms={ms:7.3f} tflops={gf/ms if ms else 0:7.1f}  {lab[:140]}")
