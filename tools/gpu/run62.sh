#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 200 python -m pytest tests/test_e2e_gpu.py -q -x -k "cli_extract or unknown_id or no_silent" 2>&1 | tail -3 | cut -c1-300
