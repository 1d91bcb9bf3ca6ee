#!/bin/bash
# GPU session 48: vae-out (decoder op list + scheduler step) vs the oracle, error paths, then the affected e2e tests.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_e2e_gpu.py -q -x -k "vae_out or unknown_id" > $O/r02_s48_vae_out.txt 2>&1
tail -30 $O/r02_s48_vae_out.txt | cut -c1-600
