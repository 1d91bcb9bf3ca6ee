#!/bin/bash
# GPU session 52: TC vs CUDA-core map kernels with bit-identical inputs (GDF_DETERMINISTIC=1), per (image, head) cosine.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
GDF_DETERMINISTIC=1 timeout 600 python tools/probe_maps_tc.py save /tmp/maps_tc.pt > $O/r02_s52_maps_tc.txt 2>&1; tail -8 $O/r02_s52_maps_tc.txt | cut -c1-220
GDF_DETERMINISTIC=1 GDF_MAPS_TC=0 timeout 600 python tools/probe_maps_tc.py cmp /tmp/maps_tc.pt > $O/r02_s52_maps_cudacore.txt 2>&1; tail -16 $O/r02_s52_maps_cudacore.txt | cut -c1-260
