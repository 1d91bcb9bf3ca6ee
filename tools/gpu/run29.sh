#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
GDF_FA_POLY8=0 timeout 200 python tools/bench_attn.py 2>&1 | head -3
GDF_FA_POLY8=9 timeout 200 python tools/bench_attn.py 2>&1 | head -3
