#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo "== deterministic"; GDF_DETERMINISTIC=1 python tools/probe_determinism_e2e2.py 2>&1 | grep -v Warn | tail -7
echo "== deterministic, plan + sync before the first call"; PROBE_PREPLAN=1 GDF_DETERMINISTIC=1 python tools/probe_determinism_e2e2.py 2>&1 | grep -v Warn | tail -7
