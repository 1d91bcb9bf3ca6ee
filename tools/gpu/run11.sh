#!/bin/bash
# GPU session 11: attention output TMA A/B, GEMM tile sweep on the small-K projections.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" > $O/r02_s11_attn_tests.txt 2>&1; tail -3 $O/r02_s11_attn_tests.txt
GDF_FA_OUT_TMA=1 timeout 200 python tools/bench_attn.py > $O/r02_s11_bench_attn_tma1.txt 2>&1; cat $O/r02_s11_bench_attn_tma1.txt
GDF_FA_OUT_TMA=0 timeout 200 python tools/bench_attn.py > $O/r02_s11_bench_attn_tma0.txt 2>&1; cat $O/r02_s11_bench_attn_tma0.txt
TUNE_LIN_ONLY=1 timeout 600 python tools/tune_gemm.py > $O/r02_s11_tune_gemm.txt 2>&1; cat $O/r02_s11_tune_gemm.txt | cut -c1-260
