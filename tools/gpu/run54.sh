#!/bin/bash
# GPU session 54: ncu --set full of the K = 1280 projection GEMMs (L2 throughput vs tensor pipe).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_tcgen05 -o $O/r02_s54_gemm1280 -f python tools/probe_gemm_l2.py > $O/r02_s54_ncu.log 2>&1; tail -4 $O/r02_s54_ncu.log
ncu -i $O/r02_s54_gemm1280.ncu-rep --page raw --csv > $O/r02_s54_gemm1280_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_s54_gemm1280_raw.csv")))
hdr = rows[0]
want = ["gpu__time_duration.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print({w: r[idx[w]] for w in want if w in idx})
print([h for h in hdr if "tensor" in h][:12])
PY
