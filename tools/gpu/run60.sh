#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02_s60.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('device %.2f img/s %.2f ms | e2e %s | clocks %s' % (d['value'], d['ms_per_step'], d['e2e'], d['clocks']['sm_mhz']))"
tail -3 gpurun_out/r02_s60.err
timeout 300 python -m pytest tests/test_e2e_gpu.py -q -x -k "tiny_xl_full_set or cli_extract" 2>&1 | tail -2
