#!/bin/bash
# Copies the evidence of the round from gpurun_out/ (scratch) to profiles/ (tracked) and derives the summaries.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out; P=profiles; T=${TAG:-r02_final}
for c in sdxl_1024 sd15_512 sd21_768_mt pixart_1024 corr_sdxl hbm_kernels reference_arm; do
  [ -s $G/${T}_bench_$c.json ] && cp $G/${T}_bench_$c.json $P/${T}_bench_$c.json
done
for c in sdxl_1024 sd15_512 sd21_768_mt pixart_1024 corr_sdxl; do
  [ -s $G/${T}_perop_$c.csv ] && cp $G/${T}_perop_$c.csv $P/${T}_perop_events_$c.csv
done
cp $G/${T}_launches_ncu.csv $P/${T}_launches_ncu.csv
cp $G/${T}_launches_summary.md $P/${T}_launches_summary.md
cp $G/${T}_hbm_launches_ncu.csv $P/${T}_hbm_launches_ncu.csv
cp $G/${T}_hbm_launches_summary.md $P/${T}_hbm_launches_summary.md
cp $G/${T}_bench_attn.txt $P/r02_bench_attn_final.txt
cp $G/${T}_bench_flux.txt $P/${T}_bench_flux.txt
cp $G/${T}_gpu_tests.txt $P/${T}_gpu_tests.txt
cp $G/r02_argmax_vs_reference_path.json $P/r02_argmax_vs_reference_path.json
cp $G/r02_full_parity_sdxl1024_b1.json $P/r02_full_parity_sdxl1024_b1.json
for f in r02_n2_bench_sdxl_1024.json r02_n2_bench_corr_sdxl.json; do [ -s $G/$f ] && cp $G/$f $P/$f; done
# ncu source-level capture of the 128-channel VAE convolution (halo mode, with residual) + per-line / per-SASS tables
cp $G/r02_s15_conv128_raw.csv $P/r02_conv128_ncu_raw.csv
cp $G/r02_s15_conv128_lines.txt $P/r02_conv128_ncu_lines.txt
python tools/ncu_sass_top.py $G/r02_s15_conv128_source.csv 40 > $P/r02_conv128_ncu_sass_top.txt 2>&1 || true
# probes quoted in DESIGN.md / the code comments
cp $G/r02_s8_probe_replan.txt $P/r02_probe_run_to_run_noise.txt
cp $G/r02_s6_trace_cross.txt $P/r02_attention_trace_cross.txt
cp $G/r02_s6_trace_self1024.txt $P/r02_attention_trace_self1024.txt
cp $G/r02_s13_conv_tests_bo0.txt $P/r02_halo_base_offset_0_tests.txt
cp $G/r02_s13_conv_tests_bo1.txt $P/r02_halo_base_offset_kx_tests.txt
for f in r02_s14_perop_halo.csv r02_s14_perop_nohalo.csv r02_s24_perop_default.csv r02_s24_perop_GDF_RES_TMA.csv r02_s43_perop_X.csv r02_s43_perop_GDF_HALO_DUAL.csv; do
  [ -s $G/$f ] && cp $G/$f $P/${f/r02_s/r02_ab_s}
done
[ -s $G/r02_s21_res_tma_probe.txt ] && cp $G/r02_s21_res_tma_probe.txt $P/r02_res_tma_probe_before_fix.txt
# DRAM traffic of the dominant kernel for bench.py's roofline.traffic
python tools/ncu_traffic.py $P/${T}_launches_ncu.csv sdxl_1024 gemm_tcgen05_kernel "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/${T}_launches_ncu.csv" $P/r02_traffic.json
python tools/sass_summary.py > $P/r02_sass_summary.md
ls -la $P | grep r02 | wc -l
