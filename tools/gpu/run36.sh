#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo "== poison, deterministic"; GDF_POISON_WORKSPACE=1 GDF_DETERMINISTIC=1 python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -7
echo "== poison, default mode"; GDF_POISON_WORKSPACE=1 python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -7
