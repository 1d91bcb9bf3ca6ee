#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1200 python -m pytest tests/test_e2e_gpu.py -q -k "flux or pixart or dit or attention_maps" 2>&1 | tail -14 | cut -c1-500
