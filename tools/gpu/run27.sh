#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_configs_gpu.py -q -k "resize or stack or corr or config2 or avgpool" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --config hbm_kernels --steps 5 --warmup 3 > $O/r02_s27_bench_hbm_kernels.json 2> $O/r02_s27_hbm.err; tail -2 $O/r02_s27_hbm.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_s27_bench_hbm_kernels.json"))
for k in d["kernels"][:6]: print("%-70s %9.1f us %7.1f GB/s %.3f write_frac %s" % (k["kernel"][:70], k["us"], k["gbs"], k["frac"], k.get("write_frac")))
PY
timeout 900 python bench.py --config sd21_768_mt --steps 5 --warmup 3 > $O/r02_s27_bench_sd21_768_mt.json 2> $O/r02_s27_sd21.err; tail -2 $O/r02_s27_sd21.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_s27_bench_sd21_768_mt.json"))
print(d["value"], d["ms_per_step"], {k: v for k, v in d["roofline"].items() if k != "tensor"})
PY
