// tcgen05 / TMEM / TMA GEMM for sm_100a, shared by every dense contraction on the extraction path:
//   linear layers (QKV, attention out-proj, FFN, proj_in/out, time-embedding MLPs),
//   3x3 convolutions as implicit GEMM (stride 1 and stride 2, zero padding by TMA out-of-bounds fill),
//   batched GEMMs (VAE single-head attention QK^T and PV).
// Reference call sites this replaces: nn.Conv2d in feature/diffusers/models/resnet.py:341,366,
// downsampling.py:147, upsampling.py:186-190; nn.Linear in attention_processor.py:3281-3289,3319,
// attention.py:1253-1257 (FeedForward), transformers/transformer_2d.py:179,207.
#pragma once
#include "ptx.cuh"

namespace gdf {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;       // 64 bf16 = 128 B = one swizzle row
constexpr int kMaxBlockN = 256;
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 4;    // TMEM accumulator ring: 2 x 256 columns, or 4 x 128 (block_n <= 128)
constexpr int kEpilogueWarps = 8;   // 4 TMEM lane quadrants x 2 interleaved column sets
constexpr int kGemmThreads = 128 + 32 * kEpilogueWarps;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps4-11 epilogue
constexpr int kStageBytesA = kBlockM * kBlockK * 2;          // 16 KB
constexpr int kStagingBytes = kEpilogueWarps * 4 * 2048;     // epilogue: per warp 4 x [32 rows][64 B] TMA-store buffers
constexpr int kBarrierBytes = 1024;                          // mbarriers + TMEM slot
constexpr int kResBytes = kEpilogueWarps * 2 * 2048;         // epilogue: per warp one 64-column residual round (TMA load)
// shared memory: [store staging | barriers | region], region = operand ring (160 KB), or, for launches that fetch the
// residual through TMA, residual staging (32 KB) + operand ring (128 KB)
constexpr int kRegionBytes = 160 * 1024;
constexpr int kGemmSmemBytes = kStagingBytes + kBarrierBytes + kRegionBytes + 1024 /*align slack*/;
static_assert(kGemmSmemBytes <= 232448, "dynamic shared memory of the GEMM kernel exceeds 227 KB");

enum GemmAMode : int { kALinear = 0, kAConvS1 = 1, kAConvS2 = 2, kAConvS1Halo = 3 };
// kAConvS1Halo: stride-1 3x3 convolution whose nine taps read ONE halo tile per 64-channel block from shared memory.
// Output tile = 16 lines x 8 pixels (128 rows); the producer loads the (18 lines x 16 pixels x 64 channels) box around
// it once (36 KB, 128B swizzle, one pixel = one 128 B row, one line = 16 rows = 2 KB) and tap (ky, kx) is the UMMA operand
// that starts (ky * 16 + kx) rows into the box with 8-row groups 2 KB apart: L2 -> smem traffic for A drops from
// 9 x 16 KB to 36 KB per channel block (the kernel is bound by that traffic at tile widths <= 160).
constexpr int kHaloLinePx = 16;                               // pixels per halo line in shared memory (10 used)
constexpr int kHaloLines = 18;
constexpr int kHaloBytes = kHaloLines * kHaloLinePx * 128;    // 36 KB per 64-channel block
constexpr int kMaxHaloStages = 4;
enum GemmAct : int { kActNone = 0, kActGeglu = 1, kActGeluTanh = 2, kActSilu = 3, kActRelu = 4 };

struct CaptureSeg {   // fp16 side output of columns [col_begin, col_end) into a feature-arena slot
  __half* ptr;        // slot base; element (row, col) goes to ptr[row * ld + (col - col_begin)]
  int col_begin, col_end, ld;
};

struct GemmParams {
  // ---- problem
  int M, N, K;          // N counts accumulator columns (for GEGLU: 2x the output width)
  int n_out;            // valid output columns (<= N, or <= N/2 for GEGLU); padding columns are dropped
  int block_n;          // multiple of 16, <= 256 (multiple of 64 when act == GEGLU)
  int num_stages;       // operand ring depth: ring bytes / (16 KB + block_n / cta_group * 128 B), <= kMaxStages
  int halo_stages;      // kAConvS1Halo: halo tiles in flight (the B tiles of the taps use num_stages)
  int halo_base_off;    // kAConvS1Halo: 1 = descriptors carry the matrix base offset of their start row (A/B knob)
  int halo_dual;        // kAConvS1Halo, single n-tile, block_n <= 128: every B (tap) tile feeds the MMAs of TWO output tiles
                        // of this CTA (two halos, two TMEM accumulators): halves the weight traffic from L2 per tile
  int cta_group;        // 1: one CTA per 128 x block_n tile; 2: CTA pair per 256 x block_n tile (tcgen05 cta_group::2)
  int num_m_tiles, num_n_tiles, num_k_blocks, batch;
  int a_mode;
  int stg_rounds;       // epilogue staging rounds in flight per warp: 2 (64 KB of staging) or 1 (32 KB, deeper ring)
  int acc_stages;       // 2 or 4 (see kMaxAccStages)
  int n_fastest;        // 1: consecutive work units walk the n-tiles of one m unit first (see decode_unit)
  int in_f16;           // 1: A and B hold fp16 (not bf16) values (feature stacks of the correspondence GEMM)
  int a_batched;        // 1: A has a batch dimension, 0: shared across the batch
  int b_batched;        // 1: B has a batch dimension, 0: shared
  // ---- K-split tail ("stream-K" for the last partial wave; linear mode, batch 1 only): work units >= sk_first are not
  // whole tiles but pieces: unit sk_first + i * sk_pieces + j = k-blocks [j * sk_kpp, (j + 1) * sk_kpp) of tile
  // sk_first + i. Every epilogue warp of a piece publishes its raw fp32 accumulators to sk_ws and counts itself in;
  // the LAST warp to arrive for a (tile, CTA rank, warp) sums the pieces in index order (deterministic), writes the sum
  // back to TMEM and runs the normal epilogue; the others just release their accumulator stage. No CTA waits for
  // another one. sk_pieces <= 1: off.
  int sk_first, sk_pieces, sk_kpp;
  float* sk_ws;                // [tail tiles][sk_pieces][cta_group][block_n columns][128 rows] fp32
  unsigned int* sk_cnt;        // [tail tiles][cta_group][kEpilogueWarps], zero between launches (the last arriver resets)
  // ---- implicit-GEMM geometry (output grid), tile = tb x th x tw pixels = 128 rows
  int B_img, H, W, tw, th, tb, tiles_x, tiles_y, cin_blocks, pad_lo;
  int tw_log2, th_log2;        // tw, th, tb are powers of two (gcd with 128)
  // ---- epilogue
  int tma_store;               // 1: bf16/fp16 destinations go through smem staging + TMA bulk stores
  int fast_epi;                // 1: tma_store and every fp32 / residual operand is 16-byte aligned (lean epilogue path)
  int res_tma;                 // 1: the lean path fetches the residual through TMA (maps.res) into smem, one round ahead
  float alpha;                 // accumulator scale (1.0 default)
  const float* bias;           // [N] column bias (already permuted for GEGLU) or null
  const float* bias_m;         // [M] row bias (transposed products) or null
  const float* row_batch_bias; // [M / rows_per_batch, N] e.g. time-embedding projection, or null
  int rows_per_batch;
  int act;
  const float* col_scale;      // [M / rows_per_batch, Nout] per-sample column gate applied before residual, or null
  const __nv_bfloat16* residual; int ld_res;   // added after activation
  int res_f16;                 // 1: residual holds fp16 bit patterns (general epilogue only: fast_epi is 0)
  float out_scale;             // multiplies (acc + residual), = 1/output_scale_factor
  __nv_bfloat16* out;  int ld_out;  long long out_batch_stride;
  int out_f16_from;            // columns >= this are written to `out` as fp16 instead of bf16 (V for fp16 P.V)
  __nv_bfloat16* out2; int ld_out2;            // second destination (skip-concat slice) or null
  float* out_f32; int ld_out_f32;              // optional fp32 destination
  __half* cap_pre; int ld_cap_pre;             // fp16 capture before residual add ("increment")
  CaptureSeg cap[3];                           // fp16 captures of the final value
  int num_cap;
  // ---- LayerNorm folded into the GEMM that consumes it (attention.py:497,525,566): A holds the un-normalised rows,
  // the weights carry gamma, bias carries beta . W^T:  value = rstd[row] * (acc - mean[row] * ln_u[col]) + bias[col]
  const float* ln_sums;        // [M][2] (sum, sum of squares) of the A rows, written by the producer's row_sums; or null
  const float* ln_u;           // [N] row sums of the (gamma-folded, bf16-rounded) weight matrix
  float ln_inv_c, ln_eps;      // 1 / row width, epsilon
  float* row_sums;             // [M][2]: this launch adds (sum, sum of squares) of its final output rows, or null
  // ---- GroupNorm statistics of the final output, for the GroupNorm that consumes it (resnet.py:327,351): per
  // (image, group) (sum, sum of squares) added with atomics. Groups of 4 / 8 / 16 channels; every 128-row tile lies
  // inside one image (checked on the host); lean epilogue path only.
  float* gn_sums;              // [images][gn_groups][2] or null
  int gn_cpg_log2, gn_groups;
  long long gn_rows_per_img;
};

// Tensor maps of one launch: operands + the six possible 16-bit destinations (TMA-store path).
struct GemmMaps {
  CUtensorMap a, b, out, out2, cap_pre, cap[3], res;
};

}  // namespace gdf
