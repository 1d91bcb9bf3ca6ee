#!/bin/bash
# GPU session 47: same-box A/B of the K-split tail and the persistent LayerNorm (order repeated to see drift).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
i=0
for v in "X=1" "GDF_STREAM_K=0" "GDF_LN_V1=1" "GDF_STREAM_K=0 GDF_LN_V1=1" "X=1" "GDF_STREAM_K=0" "GDF_STREAM_K=0 GDF_LN_V1=1"; do
  i=$((i+1))
  echo "== $v"
  env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s47_perop_$i.csv 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('%.2f img/s  %.2f ms  clocks %s  kinds %s' % (d['value'], d['ms_per_step'], d['clocks'], {k: round(v, 2) for k, v in r['per_kind_ms_per_step'].items()}))"
done
