#!/bin/bash
# GPU session 4: split-row softmax (16 softmax warps) vs thread-per-row; tests, bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention" > $O/r02_s4_attn_tests.txt 2>&1; tail -2 $O/r02_s4_attn_tests.txt
GDF_FA_SPLIT=1 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention" > $O/r02_s4_attn_tests_split1.txt 2>&1; tail -2 $O/r02_s4_attn_tests_split1.txt
for sp in 1 2; do
  for pl in 0 2 3 4; do GDF_FA_SPLIT=$sp GDF_FA_POLY8=$pl timeout 300 python tools/bench_attn.py; done
done > $O/r02_s4_bench_attn.txt 2>&1
for pl in 0 3; do GDF_FA_POLY8=$pl BENCH_ATTN_ALL=1 timeout 300 python tools/bench_attn.py | grep -v "d64"; done >> $O/r02_s4_bench_attn.txt 2>&1
cat $O/r02_s4_bench_attn.txt
GDF_FA_POLY8=3 BENCH_ATTN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o $O/r02_s4_attn python tools/bench_attn.py > $O/r02_s4_ncu.log 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_s4_gpu_tests.txt 2>&1; tail -3 $O/r02_s4_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_s4_perop.csv > $O/r02_s4_bench.json 2> $O/r02_s4_bench.err
GDF_FA_POLY8=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_s4_bench_poly3.json 2>> $O/r02_s4_bench.err
cut -c1-300 $O/r02_s4_bench.json; cut -c1-300 $O/r02_s4_bench_poly3.json
