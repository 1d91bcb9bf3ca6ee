#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
GDF_DETERMINISTIC=1 timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool initcheck --print-limit 40 --log-file gpurun_out/r02_s35_initcheck.log python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -5
grep -c "Uninitialized" gpurun_out/r02_s35_initcheck.log
grep "Uninitialized" -A6 gpurun_out/r02_s35_initcheck.log | grep -E "Uninitialized|at .*gdf|in .*\.cu|by thread" | sed 's/^=========//' | sort | uniq -c | sort -rn | head -30
