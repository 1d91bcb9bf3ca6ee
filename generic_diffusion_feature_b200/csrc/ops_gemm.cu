// Host-side construction of GEMM / implicit-GEMM launches (tensor maps + tiling).
#include "ops.h"
#include <stdlib.h>
#include <string.h>

namespace gdf {

static int gcd_int(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Tile width. Two effects compete: wide tiles need fewer operand bytes per MMA cycle (L2 -> smem ~ (128 + bn/2) / bn
// per CTA of a pair) and amortise the per-tile epilogue / handshake cost, narrow tiles can cut the padding columns of
// the last n-tile and, for the small launches of the transformer blocks (M = 8192: 32 row units), the idle tail of
// the last wave: N = 1280 is 160 work units at bn = 256 = 2.16 waves of the 74 resident CTA pairs (3 waves of 256
// columns), but 192 units at bn = 224 (3 waves of 224) or 256 units at bn = 160 (4 waves of 160).
// Cost model: waves x (bn + kTileFixedCols), kTileFixedCols = per-tile fixed cost in column equivalents.
// GDF_WAVE_POLICY=0 restores the padding-only rule (A/B timing).
// K-split of the tail wave: S pieces per tile cost ceil(nk / S) k-blocks of main loop + the fix-up: every piece writes its
// 128 x bn fp32 accumulators per CTA (~2 k-block times) and the last one reads the other S - 1 back (~3 k-block times
// each: one SM pulling 128 KB from L2). Split only when that beats the whole tile by > 15 %. GDF_SK_MAX_PIECES bounds S.
// MEASURED (profiles/r02_ksplit_ab.md): a net loss on the SDXL step (71.7 vs 73.2 images/s on the same box). The saved
// main-loop time of a short-K tail (10 of 20 k-blocks = ~3.5 us) is of the order of the fix-up's dependent latencies
// (publish 128 KB, fence, counter, read the peer's 128 KB back, TMEM store), and the pieces add 10-30 us per launch. The
// executor therefore only passes a workspace with GDF_STREAM_K=1; the op-level API splits whenever the caller gives one.
int plan_k_split(int nk, int groups, int tail, int* kpp_out, float* tail_cost) {
  int best_s = 1, best_kpp = nk;
  float best = (float)nk * 0.85f;
  const int smax = env_int("GDF_SK_MAX_PIECES", 4);
  if (tail > 0) {
    for (int s = 2; s <= smax && s * tail <= groups; ++s) {
      const int kpp = (nk + s - 1) / s;
      if (kpp < 4) break;
      const int se = (nk + kpp - 1) / kpp;   // every piece non-empty
      if (se < 2) continue;
      const float cost = (float)kpp + 2.f + 3.f * (float)(se - 1);
      if (cost < best) {
        best = cost;
        best_s = se;
        best_kpp = kpp;
      }
    }
  }
  if (best_s == 1) best = (float)nk;
  if (kpp_out) *kpp_out = best_kpp;
  if (tail_cost) *tail_cost = tail > 0 ? best : 0.f;
  return best_s;
}

int choose_block_n(int N, bool geglu, int num_m_tiles, int num_k_blocks, bool k_split) {
  const int forced = env_int("GDF_BLOCK_N", 0);   // tuning knob
  if (forced > 0 && forced % 16 == 0 && forced <= kMaxBlockN && (!geglu || forced % 64 == 0)) return forced;
  if (geglu) return N >= 256 ? 256 : 128;
  if (N <= 256) return ((N + 15) / 16) * 16;
  if (env_int("GDF_WAVE_POLICY", 1) == 0 || num_m_tiles <= 0) {
    int best = 256, best_waste = ((N + 255) / 256) * 256 - N;
    const int cands[3] = {224, 192, 160};
    for (int i = 0; i < 3; ++i) {
      const int bn = cands[i];
      const int waste = ((N + bn - 1) / bn) * bn - N;
      if (waste * 8 < best_waste * 8 - N) {   // accept a narrower tile only if it saves > 1/8 of N in padded columns
        best = bn;
        best_waste = waste;
      }
    }
    return best;
  }
  const int kTileFixedCols = 48;
  const int cg = num_m_tiles >= 2 ? 2 : 1;
  const long long m_units = (num_m_tiles + cg - 1) / cg;
  const long long resident = 148 / cg;
  int best = 256;
  long long best_cost = -1;
  // widths that divide N first: only launches without a ragged last tile qualify for the specialised ("lean" /
  // "simple") epilogue instantiations, which are worth more than the last few per cent of wave packing
  for (int pass = 0; pass < 2 && best_cost < 0; ++pass) {
    for (int bn = 256; bn >= 128; bn -= 32) {   // multiples of 32: whole rounds of the lean epilogue path
      if (pass == 0 && N % bn != 0) continue;
      const long long units = m_units * ((N + bn - 1) / bn);
      long long waves = (units + resident - 1) / resident;
      long long cost = waves * (bn + kTileFixedCols);
      if (k_split && num_k_blocks > 0 && units > resident && units % resident != 0) {
        // the last partial wave costs plan_k_split's tail instead of a whole tile (in 1/1024 of a wave)
        float tail_cost = 0.f;
        plan_k_split(num_k_blocks, (int)resident, (int)(units % resident), nullptr, &tail_cost);
        const long long milli = (units / resident) * 1024 + (long long)(1024.f * tail_cost / (float)num_k_blocks);
        cost = milli * (bn + kTileFixedCols) / 1024;
      }
      if (best_cost < 0 || cost * 100 < best_cost * 97) {   // a narrower tile has to win by > 3 %
        best = bn;
        best_cost = cost;
      }
    }
  }
  return best;
}

// CTA pair (256-row tiles, half of B per CTA) whenever the tile shape allows it; GDF_CTA_GROUP=1|2 forces.
static int choose_cta_group(int block_n, int num_m_tiles) {
  const int forced = env_int("GDF_CTA_GROUP", 0);
  const bool can = (block_n % 16 == 0) && block_n >= 32;
  if (forced == 1) return 1;
  if (forced == 2) return can ? 2 : 1;
  return (can && num_m_tiles >= 2) ? 2 : 1;
}

static void finish_tiling(GemmParams& p) {
  p.cta_group = choose_cta_group(p.block_n, p.num_m_tiles);
  // n-fastest rasterisation whenever the whole weight matrix stays L2-resident (every layer of the path: <= 26 MB)
  const long long w_bytes = (long long)p.N * p.K * 2;
  p.acc_stages = (p.block_n <= 128 && env_int("GDF_ACC4", 1) != 0) ? 4 : 2;
  p.n_fastest = (p.num_n_tiles > 1 && w_bytes <= (48ll << 20) && env_int("GDF_RASTER_N", 1) != 0) ? 1 : 0;
}
// ring depth once the epilogue mode (residual staging or not) is known
static void finish_stages(GemmParams& p) {
  // Long main loops are paced by operand latency (DRAM ~1.5 us against ~0.2 us of MMA work per stage: with 4 stages the
  // MMA issuer spent 53 % of its samples waiting on full_bar in the K = 5120 launches) and hide their epilogue anyway:
  // they run with one staging round per warp and give the 32 KB to the ring. GDF_STG1_MIN_KB tunes the threshold.
  p.stg_rounds = (p.num_k_blocks >= env_int("GDF_STG1_MIN_KB", 16)) ? 1 : 2;
  const int ring = kRegionBytes + (2 - p.stg_rounds) * (kStagingBytes / 2) - (p.res_tma ? kResBytes : 0);
  if (p.a_mode == kAConvS1Halo) {
    // two halo tiles (the one in use + the next, requested part-way through the current block: gemm_sm100.cu
    // tap_next_halo); the rest of the region is the ring of per-tap B tiles. GDF_HALO_STAGES=3 for A/B timing.
    const int b_stage = (p.block_n / p.cta_group) * kBlockK * 2;
    p.halo_stages = env_int("GDF_HALO_STAGES", 2) == 3 ? 3 : 2;
    if (ring - p.halo_stages * kHaloBytes < 3 * b_stage) p.halo_stages = 2;
    // Dual-tile mode (one n-tile of <= 128 columns, four accumulator stages): two halos per block -> four halo slots. The
    // weight tiles of a 128-channel convolution are 2x the bytes of its halos per output tile (9 x 8 KB against 36 KB
    // per channel block and CTA) and every tile re-reads them from L2; sharing each tap tile between two output tiles
    // halves that. Needs >= 4 B stages next to the 144 KB of halos: not together with the residual staging (32 KB).
    p.halo_dual = 0;
    if (p.num_n_tiles == 1 && p.block_n <= 128 && p.acc_stages == 4 && env_int("GDF_HALO_DUAL", 1) != 0 &&
        (ring - 4 * kHaloBytes) / b_stage >= 4) {
      p.halo_dual = 1;
      p.halo_stages = 4;
    }
    p.num_stages = (ring - p.halo_stages * kHaloBytes) / b_stage;
  } else {
    p.num_stages = ring / (kStageBytesA + (p.block_n / p.cta_group) * kBlockK * 2);
  }
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  if (env_int("GDF_MAX_STAGES", 0) > 0 && p.num_stages > env_int("GDF_MAX_STAGES", 0))
    p.num_stages = env_int("GDF_MAX_STAGES", 0);   // tuning knob
}

static int make_store_map(CUtensorMap* m, const GemmParams& p, const void* ptr, int width, int ld,
                          long long batch_stride) {
  if (p.a_mode == kALinear) {
    uint64_t dims[3] = {(uint64_t)width, (uint64_t)p.M, (uint64_t)p.batch};
    uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)(p.batch > 1 ? batch_stride : (long long)p.M * ld) * 2};
    uint32_t box[3] = {32, 32, 1};
    return make_tmap_bf16(m, ptr, 3, dims, str, box, 64);
  }
  const int cx = p.tw < 32 ? p.tw : 32;
  const int cy = p.th < 32 / cx ? p.th : 32 / cx;
  const int cb = 32 / (cx * cy);
  uint64_t dims[4] = {(uint64_t)width, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B_img};
  uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)p.W * ld * 2, (uint64_t)p.H * p.W * ld * 2};
  uint32_t box[4] = {32, (uint32_t)cx, (uint32_t)cy, (uint32_t)cb};
  return make_tmap_bf16(m, ptr, 4, dims, str, box, 64);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Decide whether the 16-bit destinations can go through TMA stores, and build the maps of out / out2.
static int setup_stores(GemmLaunch* g) {
  GemmParams& p = g->p;
  const int out_tile_w = (p.act == kActGeglu) ? p.block_n / 2 : p.block_n;
  bool ok = out_tile_w >= 32 && p.n_out >= 32 && p.n_out % 8 == 0;
  if (p.out && (p.ld_out % 8 != 0 || !aligned16(p.out) || (p.batch > 1 && p.out_batch_stride % 8 != 0))) ok = false;
  if (p.out2 && (p.ld_out2 % 8 != 0 || !aligned16(p.out2) || p.batch > 1)) ok = false;
  if (p.cap_pre && (p.ld_cap_pre % 8 != 0 || p.batch > 1)) ok = false;
  for (int i = 0; i < p.num_cap; ++i)
    if (p.cap[i].ld % 8 != 0 || p.cap[i].col_begin % 32 != 0 || p.batch > 1) ok = false;
  if (!p.out && !p.out2 && !p.cap_pre && p.num_cap == 0) ok = false;
  if (env_int("GDF_TMA_STORE", 1) == 0) ok = false;   // tuning knob
  p.tma_store = ok ? 1 : 0;
  p.fast_epi = 0;
  p.res_tma = 0;
  finish_stages(p);
  if (!ok) return GDF_OK;
  {
    bool f = env_int("GDF_FAST_EPI", 1) != 0;
    if (p.bias && !aligned16(p.bias)) f = false;
    if (p.row_batch_bias && (!aligned16(p.row_batch_bias) || p.N % 4 != 0)) f = false;
    if (p.col_scale && (!aligned16(p.col_scale) || p.n_out % 4 != 0)) f = false;
    if (p.residual && (!aligned16(p.residual) || p.ld_res % 8 != 0)) f = false;
    if (p.act == kActGeglu && (p.residual || p.col_scale)) f = false;   // lean GEGLU path has neither
    if (p.res_f16 || p.act == kActRelu) f = false;   // fp16 residual / ReLU exist in the general epilogue only
    if (p.ln_sums && (!aligned16(p.ln_u) || p.alpha != 1.f || p.bias_m)) f = false;
    p.fast_epi = f ? 1 : 0;
  }
  // Residual through TMA (one 64-column round per warp in flight, fetched under the main loop) for every launch of the
  // lean path: per-thread row loads after the accumulator wait put a DRAM round trip on every tile's critical path and
  // evict the bias vectors from the ~29 KB of L1 left next to 227 KB of shared memory (128-channel VAE convolution with
  // residual: 0.64 -> 0.57 ms, step +1.5 %). Round 1 kept it off for the single-staging-round launches because of
  // sparse wrong values; that was a write-after-read race in the epilogue (see the comment at the buffer reads in
  // gemm_sm100.cu), fixed in round 2. GDF_RES_TMA=0 restores the direct loads (A/B timing).
  // (dual-tile halo convolutions need the 32 KB of residual staging for their B ring: GDF_HALO_DUAL_RES=1 gives them the
  // dual mode with direct residual loads instead of the single-tile mode with the residual through TMA)
  const bool dual_wins = p.a_mode == kAConvS1Halo && p.halo_dual && env_int("GDF_HALO_DUAL_RES", 0) != 0;
  if (p.fast_epi && p.residual && p.batch == 1 && p.act != kActGeglu && env_int("GDF_RES_TMA", 1) != 0 && !dual_wins) {
    // same geometry as the store maps: [32 rows][32 columns] boxes, SWIZZLE_64B, rows / columns out of range read 0
    GDF_TRY(make_store_map(&g->maps.res, p, p.residual, p.n_out, p.ld_res, 0));
    p.res_tma = 1;
    finish_stages(p);
  }
  if (p.out) GDF_TRY(make_store_map(&g->maps.out, p, p.out, p.n_out, p.ld_out, p.out_batch_stride));
  if (p.out2) GDF_TRY(make_store_map(&g->maps.out2, p, p.out2, p.n_out, p.ld_out2, 0));
  return GDF_OK;
}

int build_capture_maps(GemmLaunch* g) {
  GemmParams& p = g->p;
  if (!p.tma_store) return GDF_OK;
  if (p.cap_pre) {
    if (!aligned16(p.cap_pre)) return fail(GDF_ERR_INVALID, "capture destination not 16-byte aligned");
    GDF_TRY(make_store_map(&g->maps.cap_pre, p, p.cap_pre, p.n_out, p.ld_cap_pre, 0));
  }
  for (int i = 0; i < p.num_cap; ++i) {
    if (!p.cap[i].ptr) continue;
    if (!aligned16(p.cap[i].ptr)) return fail(GDF_ERR_INVALID, "capture destination not 16-byte aligned");
    GDF_TRY(make_store_map(&g->maps.cap[i], p, p.cap[i].ptr, p.cap[i].col_end - p.cap[i].col_begin, p.cap[i].ld, 0));
  }
  return GDF_OK;
}

static void fill_epilogue(GemmParams& p, const Epilogue& e) {
  p.alpha = e.alpha;
  p.n_out = e.n_out > 0 ? e.n_out : (e.act == kActGeglu ? p.N / 2 : p.N);
  p.bias = e.bias;
  p.bias_m = e.bias_m;
  p.row_batch_bias = e.row_batch_bias;
  p.rows_per_batch = e.rows_per_batch;
  p.act = e.act;
  p.col_scale = e.col_scale;
  p.residual = e.residual;
  p.ld_res = e.ld_res;
  p.out_scale = e.out_scale;
  p.out = e.out;
  p.ld_out = e.ld_out;
  p.out_batch_stride = e.out_batch_stride;
  p.out_f16_from = e.out_f16_from > 0 ? e.out_f16_from : (e.out_f16_from < 0 ? 0 : (1 << 30));
  p.out2 = e.out2;
  p.ld_out2 = e.ld_out2;
  p.out_f32 = e.out_f32;
  p.ld_out_f32 = e.ld_out_f32;
  p.cap_pre = e.cap_pre;
  p.ld_cap_pre = e.ld_cap_pre;
  p.num_cap = e.num_cap;
  for (int i = 0; i < 3; ++i) p.cap[i] = e.cap[i];
  p.ln_sums = e.ln_sums;
  p.ln_u = e.ln_u;
  p.ln_inv_c = 1.f / (float)p.K;
  p.ln_eps = e.ln_eps;
  p.row_sums = e.row_sums;
  p.gn_sums = e.gn_sums;
  p.gn_cpg_log2 = e.gn_cpg == 4 ? 2 : e.gn_cpg == 8 ? 3 : 4;
  p.gn_groups = e.gn_groups;
  p.gn_rows_per_img = e.gn_rows_per_img > 0 ? e.gn_rows_per_img : 1;
  p.in_f16 = e.in_f16 ? 1 : 0;   // both builders (the convolution path used to drop it)
  p.res_f16 = (e.res_f16 && e.residual) ? 1 : 0;
}

// fused GroupNorm statistics need the lean epilogue on every column and 128-row tiles that stay inside one image
static int check_gn_stats(const GemmParams& p, const Epilogue& e) {
  if (!e.gn_sums) return GDF_OK;
  const bool cpg_ok = e.gn_cpg == 4 || e.gn_cpg == 8 || e.gn_cpg == 16;
  const bool tile_ok = (p.a_mode == kALinear) ? (e.gn_rows_per_img % kBlockM == 0) : (p.tb == 1);
  if (!cpg_ok || !tile_ok || !p.fast_epi || p.n_out % 32 != 0 || p.act == kActGeglu || p.batch != 1)
    return fail(GDF_ERR_UNSUPPORTED, "fused GroupNorm statistics: unsupported launch (cpg %d, n_out %d, fast_epi %d)",
                e.gn_cpg, p.n_out, p.fast_epi);
  return GDF_OK;
}

int build_linear(GemmLaunch* g, const bf16* A, long long M, int K, int lda, const bf16* W, int N, int ldw,
                 const Epilogue& e, int batch, long long a_batch_stride, long long w_batch_stride, int block_n) {
  memset(g, 0, sizeof(*g));
  GemmParams& p = g->p;
  const bool geglu = (e.act == kActGeglu);
  const bool sk_ok = e.sk_ws && e.sk_cnt && batch == 1;
  if (block_n <= 0)
    block_n = choose_block_n(N, geglu, (int)((M + kBlockM - 1) / kBlockM) * batch, (K + kBlockK - 1) / kBlockK, sk_ok);
  if (block_n % 16 != 0 || block_n > kMaxBlockN || (geglu && block_n % 64 != 0))
    return fail(GDF_ERR_INVALID, "build_linear: bad block_n %d", block_n);
  if (K % 8 != 0 || lda % 8 != 0 || ldw % 8 != 0)
    return fail(GDF_ERR_SHAPE, "build_linear: K/lda/ldw must be multiples of 8 (K=%d lda=%d ldw=%d)", K, lda, ldw);
  if (M <= 0 || N <= 0 || K <= 0 || M > 0x7fffffffLL) return fail(GDF_ERR_SHAPE, "build_linear: bad shape");
  p.M = (int)M;
  p.N = N;
  p.K = K;
  p.block_n = block_n;
  p.num_m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  p.num_n_tiles = (N + block_n - 1) / block_n;
  p.num_k_blocks = (K + kBlockK - 1) / kBlockK;
  p.batch = batch;
  p.a_mode = kALinear;
  p.a_batched = (batch > 1 && a_batch_stride != 0) ? 1 : 0;
  p.b_batched = (w_batch_stride != 0) ? 1 : 0;
  fill_epilogue(p, e);
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)(p.a_batched ? batch : 1)};
    uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(p.a_batched ? a_batch_stride : M * (long long)lda) * 2};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)kBlockM, 1};
    GDF_TRY(make_tmap_bf16(&g->maps.a, A, 3, dims, str, box));
  }
  finish_tiling(p);
  {
    const int wb = p.b_batched ? batch : 1;
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)wb};
    uint64_t str[2] = {(uint64_t)ldw * 2, (uint64_t)(p.b_batched ? w_batch_stride : (long long)N * ldw) * 2};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)(block_n / p.cta_group), 1};
    GDF_TRY(make_tmap_bf16(&g->maps.b, W, 3, dims, str, box));
  }
  GDF_TRY(setup_stores(g));
  GDF_TRY(check_gn_stats(p, e));
  if (sk_ok && block_n % 32 == 0) {
    // K-split of the last partial wave (GemmParams::sk_*)
    const int groups = gemm_resident_groups(p.cta_group);
    const int units = ((p.num_m_tiles + p.cta_group - 1) / p.cta_group) * p.num_n_tiles;
    const int tail = groups > 0 ? units % groups : 0;
    if (groups > 0 && units > groups && tail != 0) {
      int kpp = 0;
      const int s = plan_k_split(p.num_k_blocks, groups, tail, &kpp, nullptr);
      const long long need = (long long)tail * s * p.cta_group * kBlockM * block_n;
      if (s > 1 && need <= e.sk_ws_floats && tail * p.cta_group * kEpilogueWarps <= e.sk_cnt_len) {
        p.sk_first = units - tail;
        p.sk_pieces = s;
        p.sk_kpp = kpp;
        p.sk_ws = e.sk_ws;
        p.sk_cnt = e.sk_cnt;
      }
    }
  }
  return e.defer_capture_maps ? GDF_OK : build_capture_maps(g);
}

int build_conv3x3(GemmLaunch* g, const bf16* X, int B, int Hin, int Win, int Cin, const bf16* Wp, int N, int stride,
                  int pad_lo, const Epilogue& e, int block_n) {
  memset(g, 0, sizeof(*g));
  GemmParams& p = g->p;
  if (Cin % kBlockK != 0) return fail(GDF_ERR_SHAPE, "build_conv3x3: Cin=%d must be a multiple of 64", Cin);
  if (N % 16 != 0) return fail(GDF_ERR_SHAPE, "build_conv3x3: N=%d must be a multiple of 16 (pad the weights)", N);
  if (stride != 1 && stride != 2) return fail(GDF_ERR_UNSUPPORTED, "build_conv3x3: stride %d", stride);
  if (stride == 2 && ((Hin | Win) & 1)) return fail(GDF_ERR_SHAPE, "build_conv3x3: stride 2 needs even H, W");
  const int H = Hin / stride, W = Win / stride;  // output grid
  // halo-tile mode (gemm_sm100.cuh kAConvS1Halo): 16 x 8 pixel output tiles, one 18 x 16 pixel box per channel block.
  // GDF_CONV_HALO=0 keeps the tap-by-tap loads (A/B timing).
  bool halo = stride == 1 && pad_lo == 1 && W % 8 == 0 && H % 16 == 0 && env_int("GDF_CONV_HALO", 1) != 0;
  if (block_n <= 0) {
    const int tw0 = halo ? 8 : gcd_int(W, kBlockM), th0 = halo ? 16 : gcd_int(H, kBlockM / tw0);
    const int tb0 = kBlockM / (tw0 * th0);
    block_n = choose_block_n(N, false, (W / tw0) * (H / th0) * ((B + tb0 - 1) / tb0));
  }
  // Measured per tile width on one box (gpurun_out/r02_s14_perop_{halo,nohalo}.csv, SDXL-1024 step): 128-column tiles
  // +5..10 %, 256-column tiles +1..3 %, 160-column tiles (UNet convolutions with N = 320 / 640 / 1280) -5..8 % - those
  // keep the tap-by-tap loads. GDF_CONV_HALO=2 forces the halo mode for every width.
  if (halo && block_n > 128 && block_n < 256 && env_int("GDF_CONV_HALO", 1) != 2) halo = false;
  const int tw = halo ? 8 : gcd_int(W, kBlockM);
  const int th = halo ? 16 : gcd_int(H, kBlockM / tw);
  const int tb = kBlockM / (tw * th);
  p.M = B * H * W;
  p.N = N;
  p.K = 9 * Cin;
  p.block_n = block_n;
  p.tiles_x = W / tw;
  p.tiles_y = H / th;
  const int tiles_b = (B + tb - 1) / tb;
  p.num_m_tiles = p.tiles_x * p.tiles_y * tiles_b;
  p.num_n_tiles = (N + block_n - 1) / block_n;
  p.num_k_blocks = 9 * (Cin / kBlockK);
  p.batch = 1;
  p.a_mode = (stride == 1) ? (halo ? kAConvS1Halo : kAConvS1) : kAConvS2;
  // Measured on B200 (gpurun_out/r02_s13_conv_tests_bo{0,1}.txt): the 128B swizzle of a tcgen05 shared-memory operand
  // is a function of the ABSOLUTE address bits [7,10), exactly like the TMA write that filled the box, so a tap that
  // starts kx rows into a 1024 B pattern needs base offset 0; putting (start >> 7) & 7 into the descriptor's
  // base-offset field is applied ON TOP and gives wrong products (7 of 12 convolution tests fail).
  p.halo_base_off = env_int("GDF_HALO_BASEOFF", 0);
  p.b_batched = 0;
  p.B_img = B;
  p.H = H;
  p.W = W;
  p.tw = tw;
  p.th = th;
  p.tb = tb;
  for (p.tw_log2 = 0; (1 << p.tw_log2) < tw; ++p.tw_log2) {}
  for (p.th_log2 = 0; (1 << p.th_log2) < th; ++p.th_log2) {}
  if ((1 << p.tw_log2) != tw || (1 << p.th_log2) != th) return fail(GDF_ERR_SHAPE, "build_conv3x3: tile %d x %d", tw, th);
  p.cin_blocks = Cin / kBlockK;
  p.pad_lo = pad_lo;
  fill_epilogue(p, e);
  if (stride == 1) {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)tw, (uint32_t)th, (uint32_t)tb};
    if (halo) {
      box[1] = kHaloLinePx;   // 16 pixels per line (x0 - 1 ... x0 + 14; 10 are used): one line = 2 KB = two swizzle patterns
      box[2] = kHaloLines;    // y0 - 1 ... y0 + 16
      box[3] = 1;
    }
    GDF_TRY(make_tmap_bf16(&g->maps.a, X, 4, dims, str, box));
  } else {
    // (B, Hout, 2, Wout, 2*Cin): innermost merges (x parity, channel)
    uint64_t dims[5] = {(uint64_t)2 * Cin, (uint64_t)W, 2, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {(uint64_t)2 * Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)2 * Win * Cin * 2,
                       (uint64_t)Hin * Win * Cin * 2};
    uint32_t box[5] = {(uint32_t)kBlockK, (uint32_t)tw, 1, (uint32_t)th, (uint32_t)tb};
    GDF_TRY(make_tmap_bf16(&g->maps.a, X, 5, dims, str, box));
  }
  finish_tiling(p);
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)N, 1};
    uint64_t str[2] = {(uint64_t)p.K * 2, (uint64_t)N * p.K * 2};
    uint32_t box[3] = {(uint32_t)kBlockK, (uint32_t)(block_n / p.cta_group), 1};
    GDF_TRY(make_tmap_bf16(&g->maps.b, Wp, 3, dims, str, box));
  }
  GDF_TRY(setup_stores(g));
  GDF_TRY(check_gn_stats(p, e));
  return e.defer_capture_maps ? GDF_OK : build_capture_maps(g);
}

}  // namespace gdf
