#!/bin/bash
# GPU session 1 of round 2: diffusers probe, new attention kernel correctness + variants, full GPU tests, bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
{
  echo "== probe"; python -c "import diffusers; print('diffusers', diffusers.__version__)" 2>&1 | tail -1
  python -m pip list 2>/dev/null | grep -i -E "diffusers|transformers|accelerate|safetensors|xformers" 
  ls baseline/_ref 2>&1 | head -5; ls /opt/wheelhouse 2>/dev/null | grep -i diffus
  nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
} > $O/r02_probe.txt 2>&1
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "attention" > $O/r02_s1_attn_tests.txt 2>&1
echo "attn tests rc=$?" >> $O/r02_s1_attn_tests.txt
for v in old 0 2 3 4; do
  if [ $v = old ]; then GDF_ATTN_TC=0 timeout 300 python tools/bench_attn.py; else GDF_FA_POLY8=$v BENCH_ATTN_ALL=1 timeout 300 python tools/bench_attn.py; fi
done > $O/r02_s1_bench_attn.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02_s1_gpu_tests.txt 2>&1
echo "gpu tests rc=$?" >> $O/r02_s1_gpu_tests.txt
timeout 600 python bench.py --steps 10 --warmup 3 --profile-csv $O/r02_s1_perop.csv > $O/r02_s1_bench.json 2> $O/r02_s1_bench.err
GDF_ATTN_TC=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_s1_bench_oldattn.json 2>> $O/r02_s1_bench.err
tail -3 $O/r02_s1_attn_tests.txt; cat $O/r02_s1_bench_attn.txt; tail -3 $O/r02_s1_gpu_tests.txt; cat $O/r02_s1_bench.json | cut -c1-600
