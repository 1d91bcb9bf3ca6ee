#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for v in "X=1" "GDF_CONV_HALO=2"; do
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-csv gpurun_out/r02_s45_perop_${v%%=*}.csv 2>/dev/null | cut -c1-180
done
python tools/agg_perlaunch.py gpurun_out/r02_s45_perop_X.csv 90 | grep "unet.*mode1" | head -12
echo == forced halo
python tools/agg_perlaunch.py gpurun_out/r02_s45_perop_GDF_CONV_HALO.csv 90 | grep "unet.*mode3" | head -12
