#!/bin/bash
# GPU session 55: A/B of the dual-tile halo mode for the 128-channel VAE convolutions WITH residual (direct residual loads).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for v in "X=1" "GDF_HALO_DUAL_RES=1" "X=1" "GDF_HALO_DUAL_RES=1"; do
  echo "== $v"
  env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s55_perop_${v%%=*}.csv 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('%.2f img/s  %.2f ms  clocks %s  kinds %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v, 2) for k, v in r['per_kind_ms_per_step'].items()}))"
  python tools/agg_perlaunch.py $O/r02_s55_perop_${v%%=*}.csv 60 | grep "N=128 K=1152" | cut -c1-170
done
