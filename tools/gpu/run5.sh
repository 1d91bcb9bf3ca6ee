#!/bin/bash
# GPU session 5: ping-pong softmax groups; all bench configs; tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention" > $O/r02_s5_attn_tests.txt 2>&1; tail -2 $O/r02_s5_attn_tests.txt
GDF_FA_PP=1 GDF_FA_POLY8=3 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "attention" > $O/r02_s5_attn_tests_pp.txt 2>&1; tail -2 $O/r02_s5_attn_tests_pp.txt
for pp in 0 1; do
  for pl in 0 2 3 4; do GDF_FA_PP=$pp GDF_FA_POLY8=$pl timeout 300 python tools/bench_attn.py; done
done > $O/r02_s5_bench_attn.txt 2>&1
grep -v "Nq576" $O/r02_s5_bench_attn.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_s5_gpu_tests.txt 2>&1; tail -3 $O/r02_s5_gpu_tests.txt
for c in sd15_512 pixart_1024 sd21_768_mt corr_sdxl hbm_kernels; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 --profile-csv $O/r02_s5_perop_$c.csv > $O/r02_s5_bench_$c.json 2> $O/r02_s5_bench_$c.err
  echo "== $c rc=$?"; cut -c1-420 $O/r02_s5_bench_$c.json; tail -3 $O/r02_s5_bench_$c.err
done
