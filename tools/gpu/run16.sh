#!/bin/bash
# GPU session 17: residual streaming loads + halo policy; conv probe and step bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q > $O/r02_s17_op_tests.txt 2>&1; tail -3 $O/r02_s17_op_tests.txt
python tools/probe_conv128.py
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s17_perop.csv > $O/r02_s17_bench.json 2> $O/r02_s17_bench.err
cut -c1-330 $O/r02_s17_bench.json; tail -3 $O/r02_s17_bench.err
python tools/agg_perlaunch.py $O/r02_s17_perop.csv 40
