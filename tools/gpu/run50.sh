#!/bin/bash
# GPU session 50: compute-sanitizer over the kernels added at the end of round 2 (K-split tail, persistent LayerNorm,
# ReLU / fp16-residual epilogue, VAE decoder op list).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --print-limit 20 --log-file $O/r02_s50_memcheck.log python -m pytest tests/test_ops_gpu.py -q -x -k "k_split or layernorm or segmentor" > $O/r02_s50_memcheck_pytest.txt 2>&1
tail -3 $O/r02_s50_memcheck_pytest.txt; tail -4 $O/r02_s50_memcheck.log
timeout 900 $CS --tool memcheck --print-limit 20 --log-file $O/r02_s50_memcheck_vae.log python -m pytest tests/test_e2e_gpu.py -q -x -k "vae_out and xl" > $O/r02_s50_memcheck_vae_pytest.txt 2>&1
tail -3 $O/r02_s50_memcheck_vae_pytest.txt; tail -4 $O/r02_s50_memcheck_vae.log
timeout 600 $CS --tool synccheck --print-limit 20 --log-file $O/r02_s50_synccheck.log python -m pytest tests/test_ops_gpu.py -q -x -k "k_split or layernorm" > $O/r02_s50_synccheck_pytest.txt 2>&1
tail -3 $O/r02_s50_synccheck_pytest.txt; tail -4 $O/r02_s50_synccheck.log
