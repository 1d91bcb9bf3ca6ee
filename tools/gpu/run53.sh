#!/bin/bash
# GPU session 53: LayerNorm through a per-warp TMA ring: parity, isolated timing, same-box A/B against the register form.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "layernorm" 2>&1 | tail -3 | cut -c1-300
for v in "X=1" "GDF_LN_V2=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --config hbm_kernels --steps 20 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k in d['kernels']:
    if 'layernorm' in k['kernel']: print('  %-40s %7.1f us %7.0f GB/s' % (k['kernel'], k['us'], k['gbs']))"
done
for v in "X=1" "GDF_LN_V2=1" "X=1" "GDF_LN_V2=1"; do
  echo "== $v"
  env $v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['roofline']
print('%.2f img/s  %.2f ms  clocks %s  kinds %s' % (d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v, 2) for k, v in r['per_kind_ms_per_step'].items()}))"
done
