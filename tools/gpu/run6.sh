#!/bin/bash
# GPU session 6: attention timeline at item boundaries, cell-centric resize, sd21 bench, full tests, headline bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention or resize or correspondence" > $O/r02_s6_op_tests.txt 2>&1; tail -2 $O/r02_s6_op_tests.txt
TRACE_LIMIT=260 timeout 120 python tools/attn_trace.py 8 20 1024 77 > $O/r02_s6_trace_cross.txt 2>&1
TRACE_LIMIT=420 timeout 120 python tools/attn_trace.py 8 20 1024 1024 > $O/r02_s6_trace_self1024.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_s6_gpu_tests.txt 2>&1; tail -3 $O/r02_s6_gpu_tests.txt
for c in hbm_kernels sd21_768_mt; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > $O/r02_s6_bench_$c.json 2> $O/r02_s6_bench_$c.err
  echo "== $c rc=$?"; cut -c1-300 $O/r02_s6_bench_$c.json; tail -3 $O/r02_s6_bench_$c.err
done
GDF_RESIZE_CELL=0 timeout 900 python bench.py --config hbm_kernels --steps 5 --warmup 3 > $O/r02_s6_bench_hbm_oldresize.json 2>/dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_s6_perop.csv > $O/r02_s6_bench.json 2> $O/r02_s6_bench.err
cut -c1-300 $O/r02_s6_bench.json
