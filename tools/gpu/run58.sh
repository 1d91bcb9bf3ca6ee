#!/bin/bash
# GPU session 58: final state - whole GPU suite, smoke, one bench line.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_s58_gpu_tests.txt 2>&1
tail -3 $O/r02_s58_gpu_tests.txt | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-330
