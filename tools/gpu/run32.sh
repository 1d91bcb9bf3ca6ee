#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -k "deterministic or tiny_xl_full" 2>&1 | tail -8 | cut -c1-600
GDF_DETERMINISTIC=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
