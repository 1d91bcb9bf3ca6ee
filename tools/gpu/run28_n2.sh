#!/bin/bash
# 2-GPU session: weak scaling of the default bench, correspondence config with the index gather, stack gather over NCCL.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --gather-stacks > $O/r02_n2_bench_sdxl_1024.json 2> $O/r02_n2_bench_sdxl_1024.err; cut -c1-300 $O/r02_n2_bench_sdxl_1024.json; tail -2 $O/r02_n2_bench_sdxl_1024.err
timeout 900 $TR bench.py --gpus 2 --config corr_sdxl --steps 5 --warmup 3 > $O/r02_n2_bench_corr_sdxl.json 2> $O/r02_n2_bench_corr_sdxl.err; cut -c1-300 $O/r02_n2_bench_corr_sdxl.json; tail -2 $O/r02_n2_bench_corr_sdxl.err
python - <<'PY'
import json
for f in ("r02_n2_bench_sdxl_1024.json", "r02_n2_bench_corr_sdxl.json"):
    try:
        d = json.load(open("gpurun_out/" + f)); print(f, d["value"], d.get("gather"))
    except Exception as e: print(f, "unreadable", e)
PY
