#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "segmentor" 2>&1 | tail -12 | cut -c1-400
