#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
echo "== deterministic"; GDF_DETERMINISTIC=1 python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -6
echo "== deterministic, conv_in unfused"; GDF_DETERMINISTIC=1 GDF_CONV_IN_FUSED=0 python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -5
echo "== deterministic, no PDL"; GDF_DETERMINISTIC=1 GDF_PDL=0 python tools/probe_determinism_e2e.py 2>&1 | grep -v Warn | tail -5
