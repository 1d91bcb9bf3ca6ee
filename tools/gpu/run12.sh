#!/bin/bash
# GPU session 12: compute-sanitizer passes over the op-level tests (memcheck, synccheck; racecheck on the GEMM / attention subset).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --print-limit 20 --log-file $O/r02_s12_memcheck.log python -m pytest tests/test_ops_gpu.py -q -x > $O/r02_s12_memcheck_pytest.txt 2>&1; tail -3 $O/r02_s12_memcheck_pytest.txt; tail -5 $O/r02_s12_memcheck.log
timeout 1200 $CS --tool synccheck --print-limit 20 --log-file $O/r02_s12_synccheck.log python -m pytest tests/test_ops_gpu.py -q -x > $O/r02_s12_synccheck_pytest.txt 2>&1; tail -3 $O/r02_s12_synccheck_pytest.txt; tail -5 $O/r02_s12_synccheck.log
timeout 1500 $CS --tool racecheck --racecheck-report analysis --print-limit 30 --log-file $O/r02_s12_racecheck.log python -m pytest tests/test_ops_gpu.py -q -x -k "linear or conv3x3 or attention or groupnorm or layernorm or resize or corr" > $O/r02_s12_racecheck_pytest.txt 2>&1; tail -3 $O/r02_s12_racecheck_pytest.txt; tail -30 $O/r02_s12_racecheck.log | cut -c1-300
