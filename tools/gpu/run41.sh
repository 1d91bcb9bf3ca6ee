#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -k "tiny_xl or tiny_21 or tiny_15 or golden or control" 2>&1 | tail -4 | cut -c1-300
for v in "X=1" "GDF_TEMB_GROUP=0"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/r02_s41_perop_${v%%=*}.csv 2>/dev/null | cut -c1-180
done
python tools/agg_perlaunch.py gpurun_out/r02_s41_perop_X.csv 80 | grep -i "k4\|total"
