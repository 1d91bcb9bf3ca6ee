#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GDF_DETERMINISTIC=1 python tools/probe_determinism.py 2>&1 | tail -16
