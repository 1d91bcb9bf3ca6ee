#!/bin/bash
# GPU session 15: ncu --set full of the VAE 128-channel convolution (halo mode), source-level stall samples.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
python tools/probe_conv128.py
GDF_CONV_HALO=0 python tools/probe_conv128.py
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 -s 3 -c 1 -o $O/r02_s15_conv128_halo -f python tools/probe_conv128.py > $O/r02_s15_ncu.log 2>&1; tail -3 $O/r02_s15_ncu.log
ncu -i $O/r02_s15_conv128_halo.ncu-rep --page raw --csv > $O/r02_s15_conv128_raw.csv 2>/dev/null
ncu -i $O/r02_s15_conv128_halo.ncu-rep --page source --csv --print-source cuda,sass > $O/r02_s15_conv128_source.csv 2>/dev/null
python tools/ncu_lines.py $O/r02_s15_conv128_source.csv 45 > $O/r02_s15_conv128_lines.txt 2>&1; head -60 $O/r02_s15_conv128_lines.txt | cut -c1-200
ls -la $O/r02_s15*
