#!/bin/bash
# GPU session 8: aggregation-head fix, re-plan probe.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "aggregation" > $O/r02_s8_op_tests.txt 2>&1; tail -3 $O/r02_s8_op_tests.txt
timeout 300 python tools/probe_replan.py > $O/r02_s8_probe_replan.txt 2>&1; tail -60 $O/r02_s8_probe_replan.txt
python -c "import diffusers" 2>&1 | tail -1
