#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 500 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_final_perop_sdxl_1024.csv > $O/r02_final_bench_sdxl_1024.json 2> $O/r02_final_bench_sdxl_1024.err
cut -c1-200 $O/r02_final_bench_sdxl_1024.json; tail -2 $O/r02_final_bench_sdxl_1024.err
