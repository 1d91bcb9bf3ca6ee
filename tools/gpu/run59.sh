#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 500 python tools/probe_vae_out.py 4 > gpurun_out/r02_s59_vae_out_full.txt 2>&1; tail -5 gpurun_out/r02_s59_vae_out_full.txt | cut -c1-300
