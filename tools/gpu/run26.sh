#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python tools/probe_write_bw.py
