#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
GDF_HALO_DUAL_RES=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "conv3x3" 2>&1 | tail -2 | cut -c1-300
python tools/probe_conv128.py
GDF_HALO_DUAL_RES=1 python tools/probe_conv128.py
for v in "X=1" "GDF_HALO_DUAL_RES=1"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-180
done
