#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GDF_DETERMINISTIC=1 PROBE_TAG=det python tools/probe_determinism_vae2.py 2>&1 | grep -v Warn | tail -1
PROBE_TAG=default python tools/probe_determinism_vae2.py 2>&1 | grep -v Warn | tail -1
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -k "deterministic or tiny or full_size or share_one or reloading" 2>&1 | tail -5 | cut -c1-400
