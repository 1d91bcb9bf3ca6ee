#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_e2e_gpu.py -q 2>&1 | tail -4 | cut -c1-300
