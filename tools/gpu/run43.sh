#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q 2>&1 | tail -3 | cut -c1-300
python tools/probe_conv128.py
GDF_HALO_DUAL=0 python tools/probe_conv128.py
for v in "X=1" "GDF_HALO_DUAL=0"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/r02_s43_perop_${v%%=*}.csv 2>/dev/null | cut -c1-180
done
python tools/agg_perlaunch.py gpurun_out/r02_s43_perop_X.csv 80 | grep "N=128 \|total"
