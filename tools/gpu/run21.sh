#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== default"; python tools/probe_res_tma.py 2>&1 | tail -9
echo "== GDF_RES_TMA_WITH_STG1=1"; GDF_RES_TMA_WITH_STG1=1 python tools/probe_res_tma.py > gpurun_out/r02_s21_res_tma_probe.txt 2>&1; cat gpurun_out/r02_s21_res_tma_probe.txt | cut -c1-600
