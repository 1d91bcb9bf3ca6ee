#!/bin/bash
# GPU session 10: conv_in rewrite, attention output through TMA, masked-chunk skip.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q > $O/r02_s10_op_tests.txt 2>&1; tail -5 $O/r02_s10_op_tests.txt
timeout 200 python tools/bench_attn.py > $O/r02_s10_bench_attn.txt 2>&1; cat $O/r02_s10_bench_attn.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_s10_perop.csv > $O/r02_s10_bench.json 2> $O/r02_s10_bench.err
cut -c1-400 $O/r02_s10_bench.json; tail -3 $O/r02_s10_bench.err
python tools/agg_perlaunch.py $O/r02_s10_perop.csv 30
