#!/bin/bash
# GPU session 49: whole GPU suite after the round's late changes (segmentor head, K-split, LayerNorm, vae-out).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_s49_gpu_tests.txt 2>&1
tail -12 $O/r02_s49_gpu_tests.txt | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
