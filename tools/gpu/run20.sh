#!/bin/bash
# GPU session 20: the evidence set of the round - whole GPU suite, every BASELINE configuration, reference arm,
# ncu launch list (time + DRAM bytes per launch) of the default bench command, HBM-kernel launch list.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02_final}
timeout 2400 python -m pytest tests -m gpu -q > $O/${T}_gpu_tests.txt 2>&1; tail -4 $O/${T}_gpu_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 --profile-csv $O/${T}_perop_sdxl_1024.csv > $O/${T}_bench_sdxl_1024.json 2> $O/${T}_bench_sdxl_1024.err
cut -c1-300 $O/${T}_bench_sdxl_1024.json; tail -2 $O/${T}_bench_sdxl_1024.err
for c in sd15_512 sd21_768_mt pixart_1024 corr_sdxl hbm_kernels; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --profile-csv $O/${T}_perop_$c.csv > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err
  echo "== $c rc=$?"; cut -c1-260 $O/${T}_bench_$c.json; tail -2 $O/${T}_bench_$c.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference_arm.json 2> $O/${T}_bench_reference_arm.err; cut -c1-300 $O/${T}_bench_reference_arm.json
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_launches_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --cuda-profiler > $O/${T}_ncu_bench.log 2>&1; tail -2 $O/${T}_ncu_bench.log | cut -c1-200
python tools/ncu_launch_summary.py $O/${T}_launches_ncu.csv > $O/${T}_launches_summary.md 2>&1; head -30 $O/${T}_launches_summary.md
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize|rownorm|layernorm|avgpool|upsample|groupnorm|corr_|qk_rmsnorm" -c 200 --csv --log-file $O/${T}_hbm_launches_ncu.csv python bench.py --config hbm_kernels --steps 2 --warmup 1 > /dev/null 2>&1
python tools/ncu_launch_summary.py $O/${T}_hbm_launches_ncu.csv > $O/${T}_hbm_launches_summary.md 2>&1; cat $O/${T}_hbm_launches_summary.md
timeout 300 python tools/bench_attn.py > $O/${T}_bench_attn.txt 2>&1; BENCH_ATTN_ALL=1 timeout 300 python tools/bench_attn.py >> $O/${T}_bench_attn.txt 2>&1; tail -9 $O/${T}_bench_attn.txt
timeout 600 python tools/bench_flux.py > $O/${T}_bench_flux.txt 2>&1; tail -5 $O/${T}_bench_flux.txt
