#!/bin/bash
# GPU session 51: tensor-core attention-probability maps: golden tests, then SDXL-size timing and TC vs CUDA-core parity.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -x -k "maps or golden" > $O/r02_s51_map_tests.txt 2>&1
tail -12 $O/r02_s51_map_tests.txt | cut -c1-500
timeout 600 python tools/probe_maps_tc.py save /tmp/maps_tc.pt > $O/r02_s51_maps_tc.txt 2>&1; tail -12 $O/r02_s51_maps_tc.txt | cut -c1-220
GDF_MAPS_TC=0 timeout 600 python tools/probe_maps_tc.py cmp /tmp/maps_tc.pt > $O/r02_s51_maps_cudacore.txt 2>&1; tail -16 $O/r02_s51_maps_cudacore.txt | cut -c1-220
