#!/bin/bash
# GPU session 13: halo-tile convolution - which base-offset rule is right, then timing.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for bo in 1 0; do
  echo "== GDF_HALO_BASEOFF=$bo"
  GDF_HALO_BASEOFF=$bo timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "conv3x3" > $O/r02_s13_conv_tests_bo$bo.txt 2>&1; tail -8 $O/r02_s13_conv_tests_bo$bo.txt | cut -c1-300
done
echo "== GDF_CONV_HALO=0 (control)"
GDF_CONV_HALO=0 timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "conv3x3" 2>&1 | tail -2
