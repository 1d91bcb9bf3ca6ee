#!/bin/bash
# GPU session 9: the whole GPU suite without -x (every failure listed).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_s9_gpu_tests.txt 2>&1; tail -15 $O/r02_s9_gpu_tests.txt
