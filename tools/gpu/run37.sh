#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for v in "X=1" "GDF_CONV_HALO=0" "GDF_RES_TMA=0" "GDF_CONV_IN_FUSED=0" "GDF_EPI_LEVEL=0" "GDF_CTA_GROUP=1" "GDF_TMA_STORE=0" "GDF_CONV_HALO=0 GDF_RES_TMA=0 GDF_CONV_IN_FUSED=0"; do
  env GDF_DETERMINISTIC=1 PROBE_TAG="$v" $v python tools/probe_determinism_vae.py 2>&1 | grep -v Warn | tail -1
done
