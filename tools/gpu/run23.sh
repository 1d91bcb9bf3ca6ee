#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for shp in 2,64,640,640 1,128,320,320; do
echo "== GDF_RES_TMA_WITH_STG1=1 shape $shp"; PROBE_SHAPE=$shp PROBE_REPS=12 GDF_RES_TMA_WITH_STG1=1 python tools/probe_res_tma.py 2>&1 | grep -v " 0 wrong" | cut -c1-1500 | head -12
done
