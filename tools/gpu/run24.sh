#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for shp in 2,64,640,640 1,128,320,320; do
echo "== probe shape $shp"; PROBE_SHAPE=$shp PROBE_REPS=15 python tools/probe_res_tma.py 2>&1 | grep -c " 0 wrong"; PROBE_SHAPE=$shp PROBE_REPS=15 python tools/probe_res_tma.py 2>&1 | grep -v " 0 wrong" | cut -c1-300 | head -5
done
for i in 1 2 3 4; do timeout 300 python -m pytest tests/test_ops_gpu.py -q 2>&1 | tail -2 | cut -c1-300; done
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_configs_gpu.py -q 2>&1 | tail -4 | cut -c1-300
python tools/probe_conv128.py
for v in "default" "GDF_RES_TMA=0" "GDF_HALO_STAGES=3"; do
  if [ "$v" = "default" ]; then e=""; else e="$v"; fi
  env $e timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s24_perop_${v%%=*}.csv > $O/r02_s24_bench_${v%%=*}.json 2>/dev/null; echo "$v: $(cut -c1-200 $O/r02_s24_bench_${v%%=*}.json)"
done
python tools/agg_perlaunch.py $O/r02_s24_perop_default.csv 30
