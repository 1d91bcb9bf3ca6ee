#!/bin/bash
# GPU session 14: halo-tile convolution timing A/B, full op tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q > $O/r02_s14_op_tests.txt 2>&1; tail -3 $O/r02_s14_op_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s14_perop_halo.csv > $O/r02_s14_bench_halo.json 2> $O/r02_s14_bench.err
cut -c1-330 $O/r02_s14_bench_halo.json; tail -3 $O/r02_s14_bench.err
GDF_CONV_HALO=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s14_perop_nohalo.csv > $O/r02_s14_bench_nohalo.json 2>> $O/r02_s14_bench.err
cut -c1-330 $O/r02_s14_bench_nohalo.json
python tools/agg_perlaunch.py $O/r02_s14_perop_halo.csv 45 | grep "mode[13]\|total"
echo ==== nohalo
python tools/agg_perlaunch.py $O/r02_s14_perop_nohalo.csv 45 | grep "mode[13]\|total"
