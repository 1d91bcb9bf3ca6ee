#!/bin/bash
# GPU session 57: attention exp2 polynomial share in the power-capped step (GDF_FA_POLY8 = 0 default / 2 / 3 / 4).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "GDF_FA_POLY8=0" "GDF_FA_POLY8=2" "GDF_FA_POLY8=3" "GDF_FA_POLY8=4" "GDF_FA_POLY8=0" "GDF_FA_POLY8=3"; do
  echo "== $v"
  env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('%.2f img/s  %.2f ms  clocks %s  kinds %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v, 2) for k, v in r['per_kind_ms_per_step'].items()}))"
done
