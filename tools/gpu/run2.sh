#!/bin/bash
# GPU session 2: where does the persistent attention kernel lose time? denominator-MMA variants + ncu.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for m in 0 1 2 3; do
  for pl in 0 3; do GDF_FA_LMODE=$m GDF_FA_POLY8=$pl timeout 300 python tools/bench_attn.py; done
done > $O/r02_s2_bench_attn.txt 2>&1
cat $O/r02_s2_bench_attn.txt
GDF_FA_LMODE=0 GDF_FA_POLY8=0 BENCH_ATTN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o $O/r02_s2_attn_l0 python tools/bench_attn.py > $O/r02_s2_ncu.log 2>&1
GDF_FA_LMODE=3 GDF_FA_POLY8=0 BENCH_ATTN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o $O/r02_s2_attn_l3 python tools/bench_attn.py >> $O/r02_s2_ncu.log 2>&1
tail -5 $O/r02_s2_ncu.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_s2_gpu_tests.txt 2>&1; tail -3 $O/r02_s2_gpu_tests.txt
