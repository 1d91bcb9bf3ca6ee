#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for i in 1 2 3; do GDF_RES_TMA_WITH_STG1=1 timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "conv3x3 or linear" 2>&1 | tail -4 | cut -c1-400; done
GDF_RES_TMA_WITH_STG1=1 timeout 600 python -m pytest tests/test_e2e_gpu.py -q -x -k "tiny_xl or full_size" 2>&1 | tail -4 | cut -c1-300
GDF_RES_TMA_WITH_STG1=1 python tools/probe_conv128.py
GDF_RES_TMA_WITH_STG1=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/r02_s22_perop_restma.csv > gpurun_out/r02_s22_bench_restma.json 2>/dev/null; cut -c1-200 gpurun_out/r02_s22_bench_restma.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/r02_s22_perop_default.csv > gpurun_out/r02_s22_bench_default.json 2>/dev/null; cut -c1-200 gpurun_out/r02_s22_bench_default.json
python tools/agg_perlaunch.py gpurun_out/r02_s22_perop_restma.csv 60 | grep "res=1" | head -14
echo == default
python tools/agg_perlaunch.py gpurun_out/r02_s22_perop_default.csv 60 | grep "res=1" | head -14
