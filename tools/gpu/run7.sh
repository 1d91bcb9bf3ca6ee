#!/bin/bash
# GPU session 7: fused conv_in, aggregation head, argmax-vs-reference test, resize kernels (ncu launch list), benches.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py -x -q > $O/r02_s7_op_tests.txt 2>&1; tail -3 $O/r02_s7_op_tests.txt
GDF_FA_POLY8=0 timeout 200 python tools/bench_attn.py > $O/r02_s7_bench_attn.txt 2>&1; cat $O/r02_s7_bench_attn.txt
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_ops_gpu.py > $O/r02_s7_gpu_tests.txt 2>&1; tail -4 $O/r02_s7_gpu_tests.txt
cat $O/r02_argmax_vs_reference_path.json 2>/dev/null
python - > $O/r02_s7_copyref.txt 2>&1 <<'PY'
import torch
x = torch.randn(8, 16384, 1280, device="cuda").half()
st = torch.empty(8, 16384, 3840, dtype=torch.float16, device="cuda")
for name, fn, by in (("strided 2-D copy into a 1280-channel slice of the stack", lambda: st[:, :, :1280].copy_(x), 2 * x.numel() * 2),
                     ("contiguous copy 1 GB", lambda: st.copy_(st.roll(0)) if False else st.mul_(1.0), 2 * st.numel() * 2)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print("%s: %.1f us, %.0f GB/s" % (name, us, by / us * 1e-3))
PY
cat $O/r02_s7_copyref.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"resize|rownorm|layernorm|avgpool|upsample" -c 120 --csv --log-file $O/r02_s7_hbm_launches.csv python bench.py --config hbm_kernels > /dev/null 2>&1
for c in hbm_kernels sd21_768_mt; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > $O/r02_s7_bench_$c.json 2> $O/r02_s7_bench_$c.err
  echo "== $c rc=$?"; cut -c1-300 $O/r02_s7_bench_$c.json; tail -3 $O/r02_s7_bench_$c.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_s7_perop.csv > $O/r02_s7_bench.json 2> $O/r02_s7_bench.err
cut -c1-300 $O/r02_s7_bench.json
GDF_CONV_IN_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_s7_bench_nofuse.json 2>> $O/r02_s7_bench.err
cut -c1-300 $O/r02_s7_bench_nofuse.json
