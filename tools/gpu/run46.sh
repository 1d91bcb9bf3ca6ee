#!/bin/bash
# GPU session 46: K-split tail wave, persistent LayerNorm, segmentor head - correctness, then same-box A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "k_split" > $O/r02_s46_ksplit_tests.txt 2>&1; KS=$?
tail -15 $O/r02_s46_ksplit_tests.txt | cut -c1-400
timeout 500 python -m pytest tests/test_ops_gpu.py -q -k "not k_split" > $O/r02_s46_op_tests.txt 2>&1
tail -15 $O/r02_s46_op_tests.txt | cut -c1-400
if [ $KS -eq 0 ]; then NEW="X=1"; else NEW="GDF_STREAM_K=0"; echo "k-split tests failed: bench with GDF_STREAM_K=0"; fi
env $NEW timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s46_perop_new.csv 2>$O/r02_s46_bench_new.err | cut -c1-200
env GDF_STREAM_K=0 GDF_LN_V1=1 timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --profile-csv $O/r02_s46_perop_old.csv 2>$O/r02_s46_bench_old.err | cut -c1-200
tail -3 $O/r02_s46_bench_new.err
python tools/agg_perlaunch.py $O/r02_s46_perop_new.csv 16
echo == old
python tools/agg_perlaunch.py $O/r02_s46_perop_old.csv 16
