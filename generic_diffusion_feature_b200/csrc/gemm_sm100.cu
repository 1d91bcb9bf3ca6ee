// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution kernel (see gemm_sm100.cuh).
//
// Roles (384 threads = 12 warps, 1 CTA per SM; CTA pairs (cluster of 2, tcgen05 cta_group::2) whenever the tile allows;
// TMEM 512 columns = 2 accumulator stages of 128 lanes x 256 fp32, or 4 x 128 for tiles <= 128 columns wide):
//   warp 0    : TMA producer (one elected thread; A tile 128x64, B tile block_n / cta_group x 64, 128B swizzle,
//               num_stages-deep ring, incremental tap / channel-block coordinates for the implicit-GEMM modes)
//   warp 1    : MMA issuer   (one elected thread of the leader CTA issues tcgen05.mma 128|256 x block_n x 16 and
//               commits to the mbarriers of both CTAs)
//   warp 2    : TMEM allocator / deallocator;  warp 3: idle (holds the register budget the epilogue warps take over)
//   warps 4-11: epilogue     (4 TMEM lane quadrants x 2 interleaved column sets: tcgen05.ld 32 lanes x 32 columns ->
//               registers -> fused epilogue -> SWIZZLE_64B smem staging -> TMA bulk stores; direct 128-bit stores only
//               for destinations TMA cannot address)
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static persistent tile loop.
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace gdf {

struct TileCoord {
  int m_tile, n_tile, bz;
};
// Work unit t of the persistent loop -> (m unit, n tile, batch). The CTAs (or CTA pairs) resident at one time work
// on consecutive units. n_fastest: the n-tiles of one m unit run side by side, so the A rows are fetched from DRAM
// once and shared through L2 while the (small, L2-resident) weight matrix is streamed by every wave; m-fastest: the
// resident CTAs share one weight tile and sweep A once per n-tile, which re-reads A from DRAM whenever M x K exceeds
// L2 (measured 224 MB instead of 118 MB per launch for M=8192, N=1280, K=5120, profiles/r01_launches_ncu_s7.csv).
__device__ __forceinline__ void decode_unit(const GemmParams& p, int t, int num_m_units, int& mu, int& n_tile, int& bz) {
  // (integer division is ~35 instructions here: skipped for single-n-tile / unbatched launches)
  if (p.n_fastest) {
    const int q = (p.num_n_tiles == 1) ? t : t / p.num_n_tiles;
    n_tile = t - q * p.num_n_tiles;
    if (p.batch == 1) { mu = q; bz = 0; }
    else { bz = q / num_m_units; mu = q - bz * num_m_units; }
  } else {
    const int q = (p.batch == 1 && p.num_n_tiles == 1) ? 0 : t / num_m_units;
    mu = t - q * num_m_units;
    if (p.batch == 1) { n_tile = q; bz = 0; }
    else if (p.num_n_tiles == 1) { n_tile = 0; bz = q; }
    else { bz = q / p.num_n_tiles; n_tile = q - bz * p.num_n_tiles; }
  }
}
// K-split tail (GemmParams::sk_*): virtual unit -> (tile, k-block range). Returns the piece index, -1 for a whole tile.
__device__ __forceinline__ int decode_piece(const GemmParams& p, int& t, int& kb0, int& kb1) {
  kb0 = 0;
  kb1 = p.num_k_blocks;
  if (p.sk_pieces <= 1 || t < p.sk_first) return -1;
  const int q = t - p.sk_first;
  const int ti = q / p.sk_pieces;
  const int piece = q - ti * p.sk_pieces;
  t = p.sk_first + ti;
  kb0 = piece * p.sk_kpp;
  kb1 = min(p.num_k_blocks, kb0 + p.sk_kpp);
  return piece;
}
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int t) {
  TileCoord c;
  c.m_tile = t % p.num_m_tiles;
  int r = t / p.num_m_tiles;
  c.n_tile = r % p.num_n_tiles;
  c.bz = r / p.num_n_tiles;
  return c;
}

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float* v) {
  uint4 a, b;
  a.x = pack_bf16x2(v[0], v[1]);  a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]);  a.w = pack_bf16x2(v[6], v[7]);
  b.x = pack_bf16x2(v[8], v[9]);  b.y = pack_bf16x2(v[10], v[11]);
  b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}
__device__ __forceinline__ void store16_f16(__half* dst, const float* v) {
  uint4 a, b;
  a.x = pack_f16x2(v[0], v[1]);  a.y = pack_f16x2(v[2], v[3]);
  a.z = pack_f16x2(v[4], v[5]);  a.w = pack_f16x2(v[6], v[7]);
  b.x = pack_f16x2(v[8], v[9]);  b.y = pack_f16x2(v[10], v[11]);
  b.z = pack_f16x2(v[12], v[13]); b.w = pack_f16x2(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}

// ---------------------------------------------------------------------------------- shared epilogue math
// Loads 32 accumulator columns starting at TMEM column `c` of this tile and applies alpha, biases, activation.
__device__ __forceinline__ void load_activate32(const GemmParams& p, uint32_t taddr, int c, int out_tile_w,
                                                int n_tile, float bm, bool row_ok, int bidx, float ln_a, float ln_b,
                                                float* v) {
  uint32_t raw[32];
  tmem_ld_32x32(taddr + c, raw);
  tmem_ld_wait();
  const int acol0 = n_tile * p.block_n + c;  // accumulator column (bias index)
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = __uint_as_float(raw[j]) * p.alpha + bm;
    if (p.ln_sums) x = (acol0 + j < p.N) ? fmaf(ln_b, __ldg(p.ln_u + acol0 + j), x * ln_a) : 0.f;   // folded LayerNorm
    if (p.bias && acol0 + j < p.N) x += __ldg(p.bias + acol0 + j);
    v[j] = x;
  }
  if (p.row_batch_bias && row_ok) {
    const float* rb = p.row_batch_bias + (long long)bidx * p.N + acol0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (acol0 + j < p.N) v[j] += __ldg(rb + j);
  }
  if (p.act == kActGeglu) {
    uint32_t graw[32];
    tmem_ld_32x32(taddr + out_tile_w + c, graw);
    tmem_ld_wait();
    const int gcol0 = acol0 + out_tile_w;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float g = __uint_as_float(graw[j]) * p.alpha;
      if (p.ln_sums) g = (gcol0 + j < p.N) ? fmaf(ln_b, __ldg(p.ln_u + gcol0 + j), g * ln_a) : 0.f;
      if (p.bias && gcol0 + j < p.N) g += __ldg(p.bias + gcol0 + j);
      v[j] *= gelu_erf_f(g);
    }
  } else if (p.act == kActGeluTanh) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
  } else if (p.act == kActSilu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
  } else if (p.act == kActRelu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// column gate + residual + output scale on 32 columns of one row (direct global reads of the residual)
__device__ __forceinline__ void gate_residual32(const GemmParams& p, float* v, long long row, int bidx, int ocol0,
                                                int ncols_total, int lim, bool row_ok) {
  if (p.col_scale && row_ok) {
    const float* g = p.col_scale + (long long)bidx * ncols_total + ocol0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (ocol0 + j < lim) v[j] *= __ldg(g + j);
  }
  if (p.residual && row_ok) {
    const __nv_bfloat16* r = p.residual + row * p.ld_res + ocol0;
    if (p.res_f16) {   // fp16 residual (captured feature maps added back by the downstream ResBlock heads)
      const __half* rh = reinterpret_cast<const __half*>(r);
      if (ocol0 + 32 <= lim && (p.ld_res % 8 == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(rh) + q);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
            v[q * 8 + 2 * k] += f.x;
            v[q * 8 + 2 * k + 1] += f.y;
          }
        }
      } else {
        for (int j = 0; j < 32; ++j)
          if (ocol0 + j < lim) v[j] += __half2float(rh[j]);
      }
    } else if (ocol0 + 32 <= lim && (p.ld_res % 8 == 0)) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(r) + q);
        float2 f;
        f = unpack_bf16x2(u.x); v[q * 8 + 0] += f.x; v[q * 8 + 1] += f.y;
        f = unpack_bf16x2(u.y); v[q * 8 + 2] += f.x; v[q * 8 + 3] += f.y;
        f = unpack_bf16x2(u.z); v[q * 8 + 4] += f.x; v[q * 8 + 5] += f.y;
        f = unpack_bf16x2(u.w); v[q * 8 + 6] += f.x; v[q * 8 + 7] += f.y;
      }
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < lim) v[j] += __bfloat162float(r[j]);
    }
  }
  if (p.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
  }
}

// Direct (non-TMA) stores of one 32-column chunk of one row to every destination.
__device__ __forceinline__ void direct_store32(const GemmParams& p, const float* vpre, const float* v, long long row,
                                               int ocol0, int lim, long long out_batch_off, bool skip16) {
  const bool full = (ocol0 + 32 <= lim);
  if (p.cap_pre && !skip16) {
    __half* dst = p.cap_pre + row * p.ld_cap_pre + ocol0;
    if (full && (p.ld_cap_pre % 8 == 0)) {
      store16_f16(dst, vpre);
      store16_f16(dst + 16, vpre + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < lim) dst[j] = __float2half_rn(vpre[j]);
    }
  }
  if (p.out && !skip16 && ocol0 >= p.out_f16_from) {
    __half* dst = reinterpret_cast<__half*>(p.out + out_batch_off + row * p.ld_out + ocol0);
    for (int j = 0; j < 32; ++j)
      if (ocol0 + j < lim) dst[j] = __float2half_rn(v[j]);
  } else if (p.out && !skip16) {
    __nv_bfloat16* dst = p.out + out_batch_off + row * p.ld_out + ocol0;
    if (full && (p.ld_out % 8 == 0)) {
      store16_bf16(dst, v);
      store16_bf16(dst + 16, v + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < lim) dst[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (p.out2 && !skip16) {
    __nv_bfloat16* dst = p.out2 + row * p.ld_out2 + ocol0;
    if (full && (p.ld_out2 % 8 == 0)) {
      store16_bf16(dst, v);
      store16_bf16(dst + 16, v + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < lim) dst[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (p.out_f32) {   // (batched launches: out_batch_stride counts fp32 elements here)
    float* dst = p.out_f32 + out_batch_off + row * p.ld_out_f32 + ocol0;
    if (full && (p.ld_out_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < lim) dst[j] = v[j];
    }
  }
  if (!skip16) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      if (s < p.num_cap) {
        const CaptureSeg& cs = p.cap[s];
        if (cs.ptr && ocol0 >= cs.col_begin && ocol0 < cs.col_end) {
          __half* dst = cs.ptr + row * cs.ld + (ocol0 - cs.col_begin);
          if (ocol0 + 32 <= cs.col_end && (cs.ld % 8 == 0)) {
            store16_f16(dst, v);
            store16_f16(dst + 16, v + 16);
          } else {
            for (int j = 0; j < 32; ++j)
              if (ocol0 + j < cs.col_end) dst[j] = __float2half_rn(v[j]);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------- TMA-store staging
struct StoreCoord {   // origin of this warp's 32-row slice in the destination tensor
  int c1, c2, c3;     // linear: (row0, batch, -) ; conv: (x, y, b)
  bool conv;
};
// One warp stages its 32 rows x 32 columns (16-bit, 64 B per row, SWIZZLE_64B) and issues one TMA bulk store.
template <bool kF16>
__device__ __forceinline__ void stage_and_store(const CUtensorMap* map, uint8_t* stg_warp, int& toggle, int lane,
                                                const float* v /*32*/, int col, const StoreCoord& sc) {
  (void)toggle;
  uint8_t* buf = stg_warp;              // generic path: one buffer, fully drained around every store
  // elect.sync picks the same leader lane for the same member mask every time (PTX ISA), so issue / commit / wait
  // stay on one thread; and the compiler knows a single lane is active -> no uniform-register waterfall loops
  if (elect_one()) bulk_wait_read<0>();
  __syncwarp();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    if (kF16) {
      u.x = pack_f16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_f16x2(v[c * 8 + 2], v[c * 8 + 3]);
      u.z = pack_f16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_f16x2(v[c * 8 + 6], v[c * 8 + 7]);
    } else {
      u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
      u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    }
    // 64-byte swizzle: 16 B chunk index XOR address bits [7:9) -> (row >> 1) & 3 for 64 B rows
    *reinterpret_cast<uint4*>(buf + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = u;
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one()) {
    if (sc.conv) tma_store_4d(map, buf, col, sc.c1, sc.c2, sc.c3);
    else tma_store_3d(map, buf, col, sc.c1, sc.c2);
    bulk_commit();
    bulk_wait_read<0>();
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------- lean epilogue (fast path)
// The epilogue is instruction-issue bound at small K (profiles/r01_s3_gemm_epilogue.md): the fast path handles full,
// 16-byte-aligned 32-column chunks with vector loads and no per-element predicates; everything else takes the
// generic path above.
__device__ __forceinline__ void ldg_f32x32(const float* src, float* dst) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
    dst[q * 4 + 0] = f.x; dst[q * 4 + 1] = f.y; dst[q * 4 + 2] = f.z; dst[q * 4 + 3] = f.w;
  }
}
__device__ __forceinline__ void add_f32x32(const float* src, float* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = ldg_keep_f4(reinterpret_cast<const float4*>(src) + q);
    v[q * 4 + 0] += f.x; v[q * 4 + 1] += f.y; v[q * 4 + 2] += f.z; v[q * 4 + 3] += f.w;
  }
}
__device__ __forceinline__ void fma_f32x32(const float* src, float s, float* v) {   // v += s * src
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
    v[q * 4 + 0] = fmaf(s, f.x, v[q * 4 + 0]); v[q * 4 + 1] = fmaf(s, f.y, v[q * 4 + 1]);
    v[q * 4 + 2] = fmaf(s, f.z, v[q * 4 + 2]); v[q * 4 + 3] = fmaf(s, f.w, v[q * 4 + 3]);
  }
}
__device__ __forceinline__ void mul_f32x32(const float* src, float* v) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src) + q);
    v[q * 4 + 0] *= f.x; v[q * 4 + 1] *= f.y; v[q * 4 + 2] *= f.z; v[q * 4 + 3] *= f.w;
  }
}
// Phi(g) * g with the Abramowitz-Stegun 7.1.26 erfc tail: gelu_erf(g) = g * (g >= 0 ? 1 - h : h),
// h = 0.5 * poly(t) * t * exp(-g^2 / 2), t = 1 / (1 + 0.3275911 |g| / sqrt(2)); |abs err| <= 1e-7.
__device__ __forceinline__ float gelu_erf_lean(float g) {
  const float ag = fabsf(g);
  const float t = __frcp_rn(fmaf(0.3275911f * 0.70710678f, ag, 1.f));
  float q = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  q = fmaf(q, t, 0.5f * 1.421413741f);
  q = fmaf(q, t, 0.5f * -0.284496736f);
  q = fmaf(q, t, 0.5f * 0.254829592f);
  const float e = ex2_approx(g * g * (-0.5f * 1.4426950408889634f));
  const float h = q * t * e;
  return g * (g >= 0.f ? 1.f - h : h);
}

// GEGLU on a pair of columns, packed fp32 math: v *= g * Phi(g), Phi(g) = 0.5 + sgn(g) * (0.5 - h(|g|)) with the same
// erfc tail as gelu_erf_lean. ~10 issue slots per element instead of ~25 (the IEEE __frcp_rn alone cost a branch and
// a slow-path call per element); this epilogue is co-critical with the MMAs of the K = C GEGLU projection.
__device__ __forceinline__ void geglu_pair(float& v0, float& v1, float g0, float g1) {
  const f32x2 ag = f2_make(fabsf(g0), fabsf(g1));
  float d0, d1;
  f2_get(f2_fma(f2_splat(0.3275911f * 0.70710678f), ag, f2_splat(1.f)), d0, d1);
  const f32x2 t = f2_make(rcp_approx(d0), rcp_approx(d1));
  f32x2 q = f2_fma(f2_splat(0.5f * 1.061405429f), t, f2_splat(0.5f * -1.453152027f));
  q = f2_fma(q, t, f2_splat(0.5f * 1.421413741f));
  q = f2_fma(q, t, f2_splat(0.5f * -0.284496736f));
  q = f2_fma(q, t, f2_splat(0.5f * 0.254829592f));
  const f32x2 g = f2_make(g0, g1);
  float x0, x1;
  f2_get(f2_mul(f2_mul(g, f2_splat(-0.5f * 1.4426950408889634f)), g), x0, x1);
  const f32x2 h = f2_mul(f2_mul(q, t), f2_make(ex2_approx(x0), ex2_approx(x1)));
  float e0, e1;
  f2_get(f2_fma(h, f2_splat(-1.f), f2_splat(0.5f)), e0, e1);   // 0.5 - h >= 0
  e0 = __uint_as_float(__float_as_uint(e0) ^ (__float_as_uint(g0) & 0x80000000u));
  e1 = __uint_as_float(__float_as_uint(e1) ^ (__float_as_uint(g1) & 0x80000000u));
  const f32x2 phi = f2_add(f2_make(e0, e1), f2_splat(0.5f));
  f2_get(f2_mul(f2_mul(f2_make(v0, v1), g), phi), v0, v1);
}

template <bool kF16>
__device__ __forceinline__ void stage_chunk(uint8_t* buf, const int (&swz)[4], const float* v) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 u;
    if (kF16) {
      u.x = pack_f16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_f16x2(v[c * 8 + 2], v[c * 8 + 3]);
      u.z = pack_f16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_f16x2(v[c * 8 + 6], v[c * 8 + 7]);
    } else {
      u.x = pack_bf16x2(v[c * 8 + 0], v[c * 8 + 1]); u.y = pack_bf16x2(v[c * 8 + 2], v[c * 8 + 3]);
      u.z = pack_bf16x2(v[c * 8 + 4], v[c * 8 + 5]); u.w = pack_bf16x2(v[c * 8 + 6], v[c * 8 + 7]);
    }
    *reinterpret_cast<uint4*>(buf + swz[c]) = u;
  }
}
__device__ __forceinline__ void issue_store(const CUtensorMap* map, const uint8_t* buf, int col, const StoreCoord& sc) {
  if (sc.conv) tma_store_4d(map, buf, col, sc.c1, sc.c2, sc.c3);
  else tma_store_3d(map, buf, col, sc.c1, sc.c2);
}
// One round of the lean path: stage 32 rows x (32 or 64) columns once (two [32 rows][64 B] SWIZZLE_64B buffers) and
// send them to up to two destinations of the same 16-bit type. Each warp owns 4 buffers = 2 rounds in flight:
// before restaging a buffer pair, every bulk group but the most recent one must have finished reading.
template <bool kF16>
__device__ __forceinline__ void store_round(const CUtensorMap* map_a, int col_a, const CUtensorMap* map_b, int col_b,
                                            bool two, uint8_t* stg_warp, int& toggle, int lane, const int (&swz)[4],
                                            const float* v0, const float* v1, const StoreCoord& sc) {
  // toggle: bit 0 = buffer pair in use, bit 1 set = this launch has ONE staging round per warp (32 KB of staging
  // instead of 64 KB; the operand ring gets the difference): the previous round must have been read out entirely
  uint8_t* buf = stg_warp + (toggle & 1) * 4096;
  if (elect_one()) {
    if (toggle & 2) bulk_wait_read<0>();
    else bulk_wait_read<1>();
  }
  __syncwarp();
  stage_chunk<kF16>(buf, swz, v0);
  if (two) stage_chunk<kF16>(buf + 2048, swz, v1);
  fence_proxy_async_smem();
  __syncwarp();
  if (elect_one()) {
    issue_store(map_a, buf, col_a, sc);
    if (two) issue_store(map_a, buf + 2048, col_a + 32, sc);
    if (map_b) {
      issue_store(map_b, buf, col_b, sc);
      if (two) issue_store(map_b, buf + 2048, col_b + 32, sc);
    }
    bulk_commit();
  }
  if (!(toggle & 2)) toggle ^= 1;
}

struct ResidualRegs { uint4 u[4]; };
__device__ __forceinline__ void residual_prefetch(ResidualRegs& r, const __nv_bfloat16* src) {
#pragma unroll
  for (int q = 0; q < 4; ++q) r.u[q] = __ldg(reinterpret_cast<const uint4*>(src) + q);
}
__device__ __forceinline__ void residual_add(const ResidualRegs& r, float* v) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float2 f;
    f = unpack_bf16x2(r.u[q].x); v[q * 8 + 0] += f.x; v[q * 8 + 1] += f.y;
    f = unpack_bf16x2(r.u[q].y); v[q * 8 + 2] += f.x; v[q * 8 + 3] += f.y;
    f = unpack_bf16x2(r.u[q].z); v[q * 8 + 4] += f.x; v[q * 8 + 5] += f.y;
    f = unpack_bf16x2(r.u[q].w); v[q * 8 + 6] += f.x; v[q * 8 + 7] += f.y;
  }
}

// ---------------------------------------------------------------------------------- fused GroupNorm statistics
// Transposing butterfly: N values per lane are summed over the 32 lanes (= 32 rows of the tile); every halving step
// exchanges half of the values, so N values cost N - 1 + (5 - log2 N) shuffles and lane l ends with value l >> (5 - log2 N).
template <int N, int O>
struct GnRed {
  static __device__ __forceinline__ void run(float* a, int lane) {
    if constexpr (N > 1) {
      constexpr int H = N / 2;
      const bool up = (lane & O) != 0;
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float send = up ? a[i] : a[i + H];
        const float keep = up ? a[i + H] : a[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
      }
      if constexpr (O > 1) GnRed<H, O / 2>::run(a, lane);
    } else {
      a[0] += __shfl_xor_sync(0xffffffffu, a[0], O);
      if constexpr (O > 1) GnRed<1, O / 2>::run(a, lane);
    }
  }
};
template <int kCpgLog2>
__device__ __forceinline__ float gn_reduce_t(const float* v, bool ok, int lane) {
  constexpr int kCpg = 1 << kCpgLog2;
  constexpr int kNg = 32 / kCpg;       // groups inside 32 columns
  constexpr int kNv = 2 * kNg;
  float a[kNv];
#pragma unroll
  for (int k = 0; k < kNg; ++k) {
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < kCpg; ++j) {
      sm += v[k * kCpg + j];
      sq = fmaf(v[k * kCpg + j], v[k * kCpg + j], sq);
    }
    a[2 * k] = ok ? sm : 0.f;
    a[2 * k + 1] = ok ? sq : 0.f;
  }
  GnRed<kNv, 16>::run(a, lane);
  return a[0];   // lane l holds value (l / (32 / kNv)) of the warp's 32 rows: (sum, sum sq) per group, interleaved
}
// v: the 32 final values of this thread's row in columns [col0, col0 + 32) (all inside n_out). Returns this lane's share
// of the warp-reduced statistics; the caller keeps a running sum over the tiles of one (image, n-tile) and issues the
// atomics once per run (gn_flush) - one atomic per value and tile made the 65536-tile VAE launches atomic-bound
// (conv_in 1.26 -> 2.17 ms with statistics in the epilogue).
__device__ __forceinline__ float gn_reduce(const GemmParams& p, const float* v, bool row_ok, int lane) {
  if (p.gn_cpg_log2 == 2) return gn_reduce_t<2>(v, row_ok, lane);
  if (p.gn_cpg_log2 == 3) return gn_reduce_t<3>(v, row_ok, lane);
  return gn_reduce_t<4>(v, row_ok, lane);
}
__device__ __forceinline__ void gn_flush(const GemmParams& p, float val, long long img, int col0, int lane) {
  const int nv = 2 * (32 >> p.gn_cpg_log2);
  const int lanes_per_value = 32 / nv;
  if ((lane & (lanes_per_value - 1)) == 0)
    atomicAdd(p.gn_sums + (img * p.gn_groups + (col0 >> p.gn_cpg_log2)) * 2 + lane / lanes_per_value, val);
}

// CG = 1: one CTA per 128 x block_n tile.  CG = 2: a CTA pair (cluster of 2 on one TPC) computes a 256 x block_n
// tile with tcgen05.mma.cta_group::2 — each CTA stages its own 128 A rows and HALF of the B rows, the leader (even)
// CTA issues the MMAs for both, and each CTA's TMEM receives its own 128 accumulator rows.  Operand bytes pulled
// from L2 per MMA cycle drop from (128 + bn) to (128 + bn / 2) rows, which is what bounds this kernel.
#define EPF(field, level, fixed_value) (kLevel >= (level) ? (fixed_value) : p.field)
// kLevel: compile-time promises about the epilogue (checked by launch_gemm). The general epilogue (level 0) is ~17.6 k
// SASS instructions of mostly untaken branches; on the epilogue-bound launches instruction-fetch stalls
// (`stall_no_inst`) were the top stall reason of the epilogue warps (24 % of all samples of the 128-channel VAE
// convolution, profiles/r01_ncu_gemm_roles.md).
//   level 1 "lean":   full tiles through the TMA-store path only (no ragged / direct-store path), no folded LayerNorm,
//                     no row statistics, no row bias, out_scale = 1, no fp32 destination, activation fixed
//                     by kGeglu (GEGLU or none). Captures, cap_pre, out2, column gate, per-sample row bias, residual
//                     and GroupNorm statistics stay run-time options.
//   level 2 "simple": additionally alpha = 1 and one bf16 destination only: [bias] [+ residual] [+ GroupNorm statistics] (VAE).
// kActRt (level 1 only): the element-wise activation (GELU-tanh / SiLU: PixArt, Flux FFNs) stays a run-time option; a
// separate instantiation so that the activation code does not sit in the image of the launches that have none.
template <int CG, bool kGeglu, int kLevel = 0, bool kActRt = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = p.block_n / CG;                         // B rows staged by this CTA
  const int stage_bytes = (p.a_mode == kAConvS1Halo ? 0 : kStageBytesA) + b_rows * kBlockK * 2;
  uint8_t* stg = smem;                                      // [epilogue warp][2 * stg_rounds][32 rows][64 B]
  const int stg_bytes = p.stg_rounds * (kStagingBytes / 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stg_bytes);
  uint8_t* res_stg = smem + stg_bytes + kBarrierBytes;      // [epilogue warp][2 halves][32 rows][64 B] (res_tma only)
  smem += stg_bytes + kBarrierBytes + (p.res_tma ? kResBytes : 0);   // operand ring (takes whatever staging leaves)
  uint64_t* full_bar = bars;                                 // [kMaxStages]  (CG = 2: the leader's are used)
  uint64_t* empty_bar = bars + kMaxStages;                   // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;               // [kMaxAccStages]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + kMaxAccStages; // [kMaxAccStages]  (CG = 2: the leader's are used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2 * kMaxAccStages);
  uint64_t* res_bar = bars + 2 * kMaxStages + 2 * kMaxAccStages + 1;   // [kEpilogueWarps] residual round landed
  uint64_t* hfull_bar = res_bar + kEpilogueWarps;                      // [kMaxHaloStages] halo tile landed (kAConvS1Halo)
  uint64_t* hempty_bar = hfull_bar + kMaxHaloStages;                   // [kMaxHaloStages] its nine taps have been consumed
  const bool halo = p.a_mode == kAConvS1Halo;
  // halo mode: region = [halo ring: halo_stages x 36 KB][B ring: num_stages x b_rows x 128 B]
  uint8_t* const b_ring = smem + (halo ? p.halo_stages * kHaloBytes : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nstages = p.num_stages;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    if (p.tma_store && p.out) tma_prefetch_desc(&maps.out);
    if (p.res_tma) tma_prefetch_desc(&maps.res);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kMaxAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CG * kEpilogueWarps);
    }
    for (int i = 0; i < kEpilogueWarps; ++i) mbar_init(&res_bar[i], 1);
    for (int i = 0; i < kMaxHaloStages; ++i) {
      mbar_init(&hfull_bar[i], 1);
      mbar_init(&hempty_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_pair(tmem_slot, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlaps the previous kernel's tail; operands / destinations only from here on
  pdl_trigger();

  // work unit = CG vertically adjacent 128-row tiles x one block_n column tile
  const int num_m_units = (p.num_m_tiles + CG - 1) / CG;
  const int real_units = p.batch * num_m_units * p.num_n_tiles;
  // K-split tail: the units behind sk_first are pieces of tiles (decode_piece)
  const int total_units = p.sk_pieces > 1 ? p.sk_first + (real_units - p.sk_first) * p.sk_pieces : real_units;
  const int unit0 = blockIdx.x / CG;
  const int unit_step = gridDim.x / CG;
  const uint32_t stage_tx_bytes = (CG * kBlockM + p.block_n) * kBlockK * 2;   // bytes landing per stage, all CTAs
  // accumulator ring in TMEM (512 columns): 2 stages of 256 columns, or 4 stages of 128 for tiles <= 128 columns wide -
  // those launches (VAE 128-channel convolutions) have an epilogue about as long as their main loop, and with two
  // stages MMA(i + 2) -> epilogue(i + 2) -> MMA(i + 4) serialises through every handshake latency
  const int acc_stages = p.acc_stages;
  const uint32_t acc_stride = 512u / (uint32_t)acc_stages;

  // register rebalancing: producer / MMA / allocator warpgroup needs few registers, the two epilogue warpgroups many
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");

  if (warp == 0) {
    // ===================================================== TMA producer (every CTA loads its own A rows / B half)
    // ONE elected thread runs the whole loop (no per-k-block elect / reconvergence / warp sync) with incremental
    // coordinates (no division per k-block): an ncu source-level capture of the 1024^2 x 128-channel VAE convolution
    // showed this warp, not the tensor pipe, pacing the kernel (~760 clk per k-block against 256 clk of MMA work;
    // MMA issuer 56 % of its samples waiting on full_bar, producer 18 % on empty_bar; profiles/r01_ncu_gemm_roles.md).
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const int cinb = p.cin_blocks;
      for (int tv = unit0; !halo && tv < total_units; tv += unit_step) {
        int mu, n_tile, bz;
        int t = tv, kb_begin, kb_end;
        decode_piece(p, t, kb_begin, kb_end);
        decode_unit(p, t, num_m_units, mu, n_tile, bz);
        const int m_tile = mu * CG + (int)cta_rank;
        const int b_row0 = n_tile * p.block_n + (int)cta_rank * b_rows;
        const int bz_b = p.b_batched ? bz : 0;
        // one pipeline stage: wait for the slot, announce the bytes (leader), then the caller issues A; B follows
        auto begin_stage = [&]() -> uint8_t* {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], stage_tx_bytes);
          return smem + s * stage_bytes;
        };
        auto end_stage = [&](uint8_t* a_dst, int k0) {
          if (CG == 2) tma_load_3d_pair(a_dst + kStageBytesA, &maps.b, &full_bar[s], k0, b_row0, bz_b);
          else tma_load_3d(a_dst + kStageBytesA, &maps.b, &full_bar[s], k0, b_row0, bz_b);
          if (++s == nstages) { s = 0; ph ^= 1; }
        };
        if (p.a_mode == kALinear) {
          const int m0 = m_tile * kBlockM;
          const int bz_a = p.a_batched ? bz : 0;
          for (int kb = kb_begin, k0 = kb_begin * kBlockK; kb < kb_end; ++kb, k0 += kBlockK) {
            uint8_t* a_dst = begin_stage();
            if (CG == 2) tma_load_3d_pair(a_dst, &maps.a, &full_bar[s], k0, m0, bz_a);
            else tma_load_3d(a_dst, &maps.a, &full_bar[s], k0, m0, bz_a);
            end_stage(a_dst, k0);
          }
        } else {
          const int r = m_tile / p.tiles_x;
          const int xt = m_tile - r * p.tiles_x;
          const int bt = r / p.tiles_y;
          const int yt = r - bt * p.tiles_y;
          const int x0 = xt * p.tw;
          const int y0 = yt * p.th;
          const int b0 = bt * p.tb;           // >= B_img for the padding tile of an odd tile count: TMA zero-fills
          int k0 = 0;                         // K offset of the weight tile = (tap * cin_blocks + cb) * 64
          if (p.a_mode == kAConvS1) {
            for (int ky = 0; ky < 3; ++ky) {
              for (int kx = 0; kx < 3; ++kx) {
                for (int cb = 0, c0 = 0; cb < cinb; ++cb, c0 += kBlockK, k0 += kBlockK) {
                  uint8_t* a_dst = begin_stage();
                  if (CG == 2) tma_load_4d_pair(a_dst, &maps.a, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, b0);
                  else tma_load_4d(a_dst, &maps.a, &full_bar[s], c0, x0 + kx - 1, y0 + ky - 1, b0);
                  end_stage(a_dst, k0);
                }
              }
            }
          } else {
            // stride-2: input viewed as (B, H, 2, W, 2*Cin) with H, W the OUTPUT extents;
            // input row 2*oy + ky - pad_lo -> parity (t & 1), half-row oy + (t >> 1)
            for (int ky = 0; ky < 3; ++ky) {
              const int ty = ky - p.pad_lo;
              const int ypar = ty & 1, yoff = ty >> 1;
              for (int kx = 0; kx < 3; ++kx) {
                const int tx = kx - p.pad_lo;
                const int xpar = tx & 1, xoff = tx >> 1;
                for (int cb = 0, c0 = xpar * cinb * kBlockK; cb < cinb; ++cb, c0 += kBlockK, k0 += kBlockK) {
                  uint8_t* a_dst = begin_stage();
                  if (CG == 2) tma_load_5d_pair(a_dst, &maps.a, &full_bar[s], c0, x0 + xoff, ypar, y0 + yoff, b0);
                  else tma_load_5d(a_dst, &maps.a, &full_bar[s], c0, x0 + xoff, ypar, y0 + yoff, b0);
                  end_stage(a_dst, k0);
                }
              }
            }
          }
        }
      }
      if (halo) {
        // Halo-tile convolution: one "block" = (work unit, 64-channel block) = one 36 KB halo box + nine B tiles (taps).
        // The halo of block i + 1 is requested during the taps of block i (see tap_next_halo), so that it has most of a
        // block of MMA time (9 x 4 instructions) to arrive; the B ring is filled in tap order around it.
        const int cinb2 = p.cin_blocks;
        int hs = 0;
        uint32_t hph = 0;
        auto issue_halo = [&](int t, int cb) {
          int mu, n_tile, bz;
          decode_unit(p, t, num_m_units, mu, n_tile, bz);
          const int m_tile = mu * CG + (int)cta_rank;
          const int r = m_tile / p.tiles_x;
          const int xt = m_tile - r * p.tiles_x;
          const int bt = r / p.tiles_y;
          const int yt = r - bt * p.tiles_y;
          mbar_wait(&hempty_bar[hs], hph ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&hfull_bar[hs], (uint32_t)(CG * kHaloBytes));
          uint8_t* dst = smem + hs * kHaloBytes;
          if (CG == 2) tma_load_4d_pair(dst, &maps.a, &hfull_bar[hs], cb * kBlockK, xt * p.tw - 1, yt * p.th - 1, bt);
          else tma_load_4d(dst, &maps.a, &hfull_bar[hs], cb * kBlockK, xt * p.tw - 1, yt * p.th - 1, bt);
          if (++hs == p.halo_stages) { hs = 0; hph ^= 1; }
        };
        const uint32_t b_tx_bytes = (uint32_t)p.block_n * kBlockK * 2;
        // With two halo tiles the slot of block i + 1 is the one block i - 1 used: it is free once the MMA issuer has
        // started block i, which the producer knows when it has been granted the B slot of tap `nstages` of block i
        // (the ring is nstages deep). Asking earlier would park the producer on hempty with the B ring running dry.
        const int tap_next_halo = p.halo_stages >= 3 && !p.halo_dual ? 0 : (nstages < 8 ? nstages : 8);
        // dual mode: a "block" covers the two tiles t and t + unit_step of this CTA (two halos per channel block, one set
        // of nine B tiles); the epilogue still sees them as consecutive tiles in consecutive accumulator stages
        const int t_stride = p.halo_dual ? 2 * unit_step : unit_step;
        auto issue_block = [&](int t, int cb) {
          issue_halo(t, cb);
          if (p.halo_dual && t + unit_step < total_units) issue_halo(t + unit_step, cb);
        };
        if (unit0 < total_units) issue_block(unit0, 0);
        for (int t = unit0; t < total_units; t += t_stride) {
          int mu, n_tile, bz;
          decode_unit(p, t, num_m_units, mu, n_tile, bz);
          const int b_row0 = n_tile * p.block_n + (int)cta_rank * b_rows;
          for (int cb = 0; cb < cinb2; ++cb) {
            for (int tap = 0; tap < 9; ++tap) {
              if (tap == tap_next_halo) {
                if (cb + 1 < cinb2) issue_block(t, cb + 1);
                else if (t + t_stride < total_units) issue_block(t + t_stride, 0);
              }
              mbar_wait(&empty_bar[s], ph ^ 1);
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], b_tx_bytes);
              const int k0 = (tap * cinb2 + cb) * kBlockK;
              uint8_t* b_dst = b_ring + s * stage_bytes;
              if (CG == 2) tma_load_3d_pair(b_dst, &maps.b, &full_bar[s], k0, b_row0, 0);
              else tma_load_3d(b_dst, &maps.b, &full_bar[s], k0, b_row0, 0);
              if (++s == nstages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================== MMA issuer (leader CTA only when CG = 2), one elected thread
    if (cta_rank == 0 && elect_one()) {
      const uint32_t idesc = p.in_f16 ? umma_idesc_f16(CG * kBlockM, p.block_n) : umma_idesc_bf16(CG * kBlockM, p.block_n);
      int s = 0;
      uint32_t ph = 0;
      int as = 0;
      uint32_t aph = 0;
      int hs = 0;
      uint32_t hph = 0;
      const int last_kb = p.num_k_blocks - 1;
      if (halo && p.halo_dual) {
        // two output tiles per B tile: tiles t and t + unit_step of this CTA, accumulators `as` and the next stage
        for (int t = unit0; t < total_units; t += 2 * unit_step) {
          const bool two = t + unit_step < total_units;
          int as2 = as + 1;
          uint32_t aph2 = aph;
          if (as2 == acc_stages) { as2 = 0; aph2 ^= 1; }
          mbar_wait(&tempty_bar[as], aph ^ 1);
          if (two) mbar_wait(&tempty_bar[as2], aph2 ^ 1);
          tc_fence_after();
          const uint32_t acc0 = tmem_base + as * acc_stride, acc1 = tmem_base + as2 * acc_stride;
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            int hs2 = hs + 1;
            uint32_t hph2 = hph;
            if (hs2 == p.halo_stages) { hs2 = 0; hph2 ^= 1; }
            mbar_wait(&hfull_bar[hs], hph);
            if (two) mbar_wait(&hfull_bar[hs2], hph2);
            tc_fence_after();
            const uint32_t h0 = smem_u32(smem + hs * kHaloBytes), h1 = smem_u32(smem + hs2 * kHaloBytes);
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap - 3 * ky;
              mbar_wait(&full_bar[s], ph);
              tc_fence_after();
              const uint32_t toff = (uint32_t)(ky * kHaloLinePx + kx) * 128u;
              const uint64_t da0 = umma_desc_kmajor_sw128_sbo(h0 + toff, kHaloLinePx * 128u, 0u);
              const uint64_t da1 = umma_desc_kmajor_sw128_sbo(h1 + toff, kHaloLinePx * 128u, 0u);
              const uint64_t db = umma_desc_kmajor_sw128(smem_u32(b_ring + s * stage_bytes));
              const uint32_t accum0 = (cb | tap) != 0 ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if (CG == 2) umma_f16_ss_pair(acc0, da0 + 2 * k, db + 2 * k, idesc, (accum0 | (uint32_t)k) != 0 ? 1u : 0u);
                else umma_f16_ss(acc0, da0 + 2 * k, db + 2 * k, idesc, (accum0 | (uint32_t)k) != 0 ? 1u : 0u);
              }
              if (two) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  if (CG == 2) umma_f16_ss_pair(acc1, da1 + 2 * k, db + 2 * k, idesc, (accum0 | (uint32_t)k) != 0 ? 1u : 0u);
                  else umma_f16_ss(acc1, da1 + 2 * k, db + 2 * k, idesc, (accum0 | (uint32_t)k) != 0 ? 1u : 0u);
                }
              }
              const bool last = (cb == p.cin_blocks - 1) && tap == 8;
              if (CG == 2) {
                umma_commit_pair(&empty_bar[s], 3);
                if (tap == 8) {
                  umma_commit_pair(&hempty_bar[hs], 3);
                  if (two) umma_commit_pair(&hempty_bar[hs2], 3);
                }
                if (last) {
                  umma_commit_pair(&tfull_bar[as], 3);
                  if (two) umma_commit_pair(&tfull_bar[as2], 3);
                }
              } else {
                umma_commit(&empty_bar[s]);
                if (tap == 8) {
                  umma_commit(&hempty_bar[hs]);
                  if (two) umma_commit(&hempty_bar[hs2]);
                }
                if (last) {
                  umma_commit(&tfull_bar[as]);
                  if (two) umma_commit(&tfull_bar[as2]);
                }
              }
              if (++s == nstages) { s = 0; ph ^= 1; }
            }
            if (two) { hs = hs2; hph = hph2; }
            if (++hs == p.halo_stages) { hs = 0; hph ^= 1; }
          }
          if (two) { as = as2; aph = aph2; }
          if (++as == acc_stages) { as = 0; aph ^= 1; }
        }
      } else
      for (int t = unit0; t < total_units; t += unit_step) {
        mbar_wait(&tempty_bar[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + as * acc_stride;
        if (halo) {
          // nine taps per halo tile: tap (ky, kx) = the operand that starts (ky * 16 + kx) rows into the box, 8-row groups
          // (one line of the 8-pixel-wide output tile) 2 KB apart; its swizzle phase is the row offset kx
          for (int cb = 0; cb < p.cin_blocks; ++cb) {
            mbar_wait(&hfull_bar[hs], hph);
            tc_fence_after();
            const uint32_t h_addr = smem_u32(smem + hs * kHaloBytes);
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap - 3 * ky;
              mbar_wait(&full_bar[s], ph);
              tc_fence_after();
              const uint64_t da = umma_desc_kmajor_sw128_sbo(h_addr + (uint32_t)(ky * kHaloLinePx + kx) * 128u,
                                                             kHaloLinePx * 128u, p.halo_base_off ? (uint32_t)kx : 0u);
              const uint64_t db = umma_desc_kmajor_sw128(smem_u32(b_ring + s * stage_bytes));
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if (CG == 2) umma_f16_ss_pair(tmem_acc, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) != 0 ? 1u : 0u);
                else umma_f16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, (cb | tap | k) != 0 ? 1u : 0u);
              }
              const bool last = (cb == p.cin_blocks - 1) && tap == 8;
              if (CG == 2) {
                umma_commit_pair(&empty_bar[s], 3);
                if (tap == 8) umma_commit_pair(&hempty_bar[hs], 3);
                if (last) umma_commit_pair(&tfull_bar[as], 3);
              } else {
                umma_commit(&empty_bar[s]);
                if (tap == 8) umma_commit(&hempty_bar[hs]);
                if (last) umma_commit(&tfull_bar[as]);
              }
              if (++s == nstages) { s = 0; ph ^= 1; }
            }
            if (++hs == p.halo_stages) { hs = 0; hph ^= 1; }
          }
          if (++as == acc_stages) { as = 0; aph ^= 1; }
          continue;
        }
        int kb_begin = 0, kb_last = last_kb;
        if (p.sk_pieces > 1) {
          int tt = t, kb_end;
          decode_piece(p, tt, kb_begin, kb_end);
          kb_last = kb_end - 1;
        }
        for (int kb = kb_begin; kb <= kb_last; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint64_t da = umma_desc_kmajor_sw128(a_addr);
          const uint64_t db = umma_desc_kmajor_sw128(a_addr + kStageBytesA);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            if (CG == 2) umma_f16_ss_pair(tmem_acc, da + 2 * k, db + 2 * k, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            else umma_f16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
          }
          if (CG == 2) {
            umma_commit_pair(&empty_bar[s], 3);       // frees this stage in both CTAs
            if (kb == kb_last) umma_commit_pair(&tfull_bar[as], 3);
          } else {
            umma_commit(&empty_bar[s]);
            if (kb == kb_last) umma_commit(&tfull_bar[as]);
          }
          if (++s == nstages) { s = 0; ph ^= 1; }
        }
        if (++as == acc_stages) { as = 0; aph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================== epilogue: 8 warps = 4 TMEM lane quadrants x 2 column sets
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int e = warp - 4;
    const int ew = e & 3;     // == warp % 4: TMEM lane quadrant this warp may access
    const int cset = e >> 2;  // this warp handles 32-column chunks with (chunk index & 1) == cset
    const int r_in_tile = ew * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    constexpr bool geglu = kGeglu;   // host dispatch: p.act == kActGeglu
    const int ncols_out = p.n_out;
    const int out_tile_w = geglu ? p.block_n / 2 : p.block_n;
    uint8_t* stg_warp = stg + e * (4096 * p.stg_rounds);   // 2 staging buffers of 2 KB per round and warp
    int toggle = p.stg_rounds == 1 ? 2 : 0;
    int swz[4];   // byte offsets of this lane's four 16 B pieces in a [32 rows][64 B] SWIZZLE_64B staging buffer
#pragma unroll
    for (int c = 0; c < 4; ++c) swz[c] = lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4);
    // residual through TMA (res_tma): one 64-column round per warp in flight, fetched while the MMAs of the tile (or
    // the stores of the previous round) run; per-thread row-strided global loads made the epilogue latency bound
    const bool res_tma = !kGeglu && p.res_tma != 0;
    uint8_t* res_buf = res_stg + e * 4096;
    uint64_t* my_res_bar = &res_bar[e];
    uint32_t* res_sink = reinterpret_cast<uint32_t*>(bars + 64) + e;   // scratch word of this warp (see the residual reads)
    uint32_t res_ph = 0;
    bool res_pending = false;
    // GroupNorm statistics (gn_sums): running per-lane sums of this warp's columns over the tiles of one
    // (image, n-tile) run, [round c < 128 | c >= 128][32-column half]; flushed with atomics when the run ends
    float gn_run0[2] = {0.f, 0.f}, gn_run1[2] = {0.f, 0.f};
    int gn_key = -1;
    auto gn_flush_run = [&]() {
      const long long img = gn_key / p.num_n_tiles;
      const int col_base = (int)(gn_key % p.num_n_tiles) * out_tile_w + cset * 64;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (col_base + hh * 32 + 32 <= ncols_out && cset * 64 + hh * 32 < out_tile_w)
          gn_flush(p, gn_run0[hh], img, col_base + hh * 32, lane);
        if (col_base + 128 + hh * 32 + 32 <= ncols_out && cset * 64 + 128 + hh * 32 < out_tile_w)
          gn_flush(p, gn_run1[hh], img, col_base + 128 + hh * 32, lane);
        gn_run0[hh] = 0.f;
        gn_run1[hh] = 0.f;
      }
    };
    for (int tv = unit0; tv < total_units; tv += unit_step) {
      TileCoord tc;
      int t = tv;
      {
        int mu, kb_begin, kb_end;
        const int piece = decode_piece(p, t, kb_begin, kb_end);
        decode_unit(p, t, num_m_units, mu, tc.n_tile, tc.bz);
        tc.m_tile = mu * CG + (int)cta_rank;
        // (general epilogue instantiation only: launch_gemm sends K-split launches there, so that the ~2.8 k instructions of
        // the fix-up do not sit inside the tile loop of the lean / simple images, which are instruction-fetch sensitive)
        if constexpr (kLevel == 0)
        if (piece >= 0) {
          // ---- K-split tail piece: publish the raw accumulators, count in; the last warp of this (tile, rank, warp) sums
          // every piece in index order into TMEM and falls through to the normal epilogue, the others are done
          mbar_wait(&tfull_bar[as], aph);
          tc_fence_after();
          const uint32_t taddr0 = tmem_base + (uint32_t(ew * 32) << 16) + as * acc_stride;
          const int np = p.sk_pieces;
          const int ti = t - p.sk_first;
          const size_t piece_stride = (size_t)CG * kBlockM * p.block_n;
          float* wsb = p.sk_ws + ((size_t)ti * np * CG + cta_rank) * (size_t)(kBlockM * p.block_n) + r_in_tile;
          const int half_w = geglu ? p.block_n / 2 : p.block_n;   // GEGLU: value chunk c and its gate chunk half_w + c
          {
            float* dst = wsb + (size_t)piece * piece_stride;
            for (int c = cset * 64; c < half_w; c += ((c & 32) ? 96 : 32)) {
              for (int g2 = 0; g2 < (geglu ? 2 : 1); ++g2) {
                const int ac = c + g2 * half_w;
                uint32_t raw[32];
                tmem_ld_32x32(taddr0 + ac, raw);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[(size_t)(ac + j) * kBlockM] = __uint_as_float(raw[j]);
              }
            }
          }
          __threadfence();
          __syncwarp();
          unsigned int* cnt = p.sk_cnt + ((size_t)ti * CG + cta_rank) * kEpilogueWarps + e;
          unsigned int old = 0;
          if (lane == 0) old = atomicAdd(cnt, 1u);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (old != (unsigned int)(np - 1)) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&tempty_bar[as], 0);
              else mbar_arrive(&tempty_bar[as]);
            }
            if (++as == acc_stages) { as = 0; aph ^= 1; }
            continue;
          }
          __threadfence();
          if (lane == 0) *cnt = 0u;   // ready for the next launch
          for (int c = cset * 64; c < half_w; c += ((c & 32) ? 96 : 32)) {
            for (int g2 = 0; g2 < (geglu ? 2 : 1); ++g2) {
              const int ac = c + g2 * half_w;
              float acc[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) acc[j] = 0.f;
              for (int j2 = 0; j2 < np; ++j2) {   // index order, this piece's own share straight from TMEM: the sum does
                if (j2 == piece) {                 // not depend on which piece arrived last
                  uint32_t raw[32];
                  tmem_ld_32x32(taddr0 + ac, raw);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(raw[j]);
                } else {
                  const float* src = wsb + (size_t)j2 * piece_stride + (size_t)ac * kBlockM;
#pragma unroll
                  for (int j = 0; j < 32; ++j) acc[j] += __ldcg(src + (size_t)j * kBlockM);
                }
              }
              tmem_st_32x32(taddr0 + ac, *reinterpret_cast<uint32_t(*)[32]>(acc));
            }
          }
          tmem_st_wait();
        }
      }
      // ---- row of this thread, origin of this warp's 32-row slice
      long long row;
      bool row_ok;
      StoreCoord sc;
      if (p.a_mode == kALinear) {
        row = (long long)tc.m_tile * kBlockM + r_in_tile;
        row_ok = row < p.M;
        sc.conv = false;
        sc.c1 = tc.m_tile * kBlockM + ew * 32;
        sc.c2 = tc.bz;
        sc.c3 = 0;
      } else {
        const int r = tc.m_tile / p.tiles_x;
        const int xt = tc.m_tile - r * p.tiles_x;
        const int bt = r / p.tiles_y;
        const int yt = r - bt * p.tiles_y;
        // tw, th, tb are powers of two (gcd with 128): shifts / masks instead of divisions
        const int tx = r_in_tile & (p.tw - 1);
        const int r2 = r_in_tile >> p.tw_log2;
        const int ty = r2 & (p.th - 1);
        const int tbi = r2 >> p.th_log2;
        const int b = bt * p.tb + tbi;
        row = ((long long)b * p.H + (yt * p.th + ty)) * p.W + (xt * p.tw + tx);
        row_ok = b < p.B_img;
        const int w0 = ew * 32;   // first row of the warp slice, decomposed the same way
        sc.conv = true;
        sc.c1 = xt * p.tw + (w0 & (p.tw - 1));
        sc.c2 = yt * p.th + ((w0 >> p.tw_log2) & (p.th - 1));
        sc.c3 = bt * p.tb + ((w0 >> p.tw_log2) >> p.th_log2);
      }
      const int bidx = (p.rows_per_batch > 0 && row_ok) ? (int)(row / p.rows_per_batch) : 0;
      if (p.gn_sums) {   // (row validity and the image are uniform over a tile: checked on the host)
        const int key = row_ok ? (int)(row / p.gn_rows_per_img) * p.num_n_tiles + tc.n_tile : -1;
        if (key != gn_key) {
          if (gn_key >= 0) gn_flush_run();
          gn_key = key;
        }
      }
      const long long out_batch_off = (long long)tc.bz * p.out_batch_stride;
      const float bm = (EPF(bias_m, 1, nullptr) && row_ok) ? __ldg(EPF(bias_m, 1, nullptr) + row) : 0.f;
      // folded LayerNorm of the A rows: value = ln_a * acc + ln_b * u[col] + bias[col]
      float ln_a = 1.f, ln_b = 0.f;
      if (EPF(ln_sums, 1, nullptr) && row_ok) {
        const float2 sq = __ldg(reinterpret_cast<const float2*>(EPF(ln_sums, 1, nullptr)) + row);
        const float mean = sq.x * p.ln_inv_c;
        const float var = fmaxf(fmaf(sq.y, p.ln_inv_c, -mean * mean), 0.f);
        ln_a = rsqrtf(var + p.ln_eps);
        ln_b = -ln_a * mean;
      }
      float rs_sum = 0.f, rs_sq = 0.f;   // row statistics of this thread's final values (row_sums)

      // per-row operand bases of the lean path (null rows of a padding tile read nothing)
      const float* rbb = (EPF(row_batch_bias, 2, nullptr) && row_ok) ? EPF(row_batch_bias, 2, nullptr) + (long long)bidx * p.N : nullptr;
      const float* csrow = (!kGeglu && EPF(col_scale, 2, nullptr) && row_ok) ? EPF(col_scale, 2, nullptr) + (long long)bidx * ncols_out : nullptr;
      // (GEGLU launches carry no residual / column gate on the lean path: checked on the host)
      const __nv_bfloat16* resrow = (!kGeglu && p.residual && row_ok && !res_tma) ? p.residual + row * p.ld_res : nullptr;
      // a round of the lean path = 64 accumulator columns starting at c (32 at the ragged end of the tile)
      auto lean_round = [&](int c) -> bool {
        if (!EPF(fast_epi, 1, 1) || c >= out_tile_w) return false;
        const int oc = tc.n_tile * out_tile_w + c;
        if (oc >= ncols_out) return false;
        const bool tw2 = (c + 32 < out_tile_w) && (oc + 32 < ncols_out);
        const int wc = tw2 ? 64 : 32;
        return oc + wc <= ncols_out && tc.n_tile * p.block_n + c + (kGeglu ? out_tile_w : 0) + wc <= p.N;
      };
      auto res_issue = [&](int c) {   // both 32-column halves; out-of-range rows / columns arrive as zeros
        const int oc = tc.n_tile * out_tile_w + c;
        if (elect_one()) {
          mbar_arrive_expect_tx(my_res_bar, 4096);
          if (sc.conv) {
            tma_load_4d(res_buf, &maps.res, my_res_bar, oc, sc.c1, sc.c2, sc.c3);
            tma_load_4d(res_buf + 2048, &maps.res, my_res_bar, oc + 32, sc.c1, sc.c2, sc.c3);
          } else {
            tma_load_3d(res_buf, &maps.res, my_res_bar, oc, sc.c1, sc.c2);
            tma_load_3d(res_buf + 2048, &maps.res, my_res_bar, oc + 32, sc.c1, sc.c2);
          }
        }
        __syncwarp();
        res_pending = true;
      };
      if (res_tma && !res_pending && lean_round(cset * 64)) res_issue(cset * 64);   // lands during the main loop
      ResidualRegs rr[2];
#pragma unroll
      for (int q = 0; q < 4; ++q) rr[0].u[q] = rr[1].u[q] = make_uint4(0u, 0u, 0u, 0u);
      // Direct residual loads (launches without res_tma): the first round's rows are requested BEFORE the accumulator
      // wait, so that their DRAM latency runs under the main loop. Issued after it, the latency sat on the critical path
      // of every tile whose warps have a single round (N <= 128: the 128-channel VAE convolution with residual ran
      // 0.70 ms against 0.53 ms without, 0.64 ms with this; ncu showed the wait on the first instructions that consume a
      // global load). Requesting the NEXT tile's rows at the end of a tile measured slower (0.69 ms: the extra tile
      // geometry and 32 more live registers cost more than the latency they hide).
      bool have_res = false;
      if (resrow && lean_round(cset * 64)) {
        const int oc = tc.n_tile * out_tile_w + cset * 64;
        residual_prefetch(rr[0], resrow + oc);
        if ((cset * 64 + 32 < out_tile_w) && (oc + 32 < ncols_out)) residual_prefetch(rr[1], resrow + oc + 32);
        have_res = true;
      }

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * acc_stride;

      // ---- lean path: rounds of 64 columns (two 32-column halves), TMEM loads / bias vectors / residual prefetch
      // in flight together, one fence + one elected TMA issue per destination and round
      int c_done = cset * 64;    // first column this warp still has to handle on the generic path
      if (EPF(fast_epi, 1, 1)) {
        for (int c = cset * 64; c < out_tile_w; c += 128) {
          const int ocol0 = tc.n_tile * out_tile_w + c;
          if (ocol0 >= ncols_out) { c_done = out_tile_w; break; }
          const int acol0 = tc.n_tile * p.block_n + c;  // accumulator column (bias index)
          const int gofs = geglu ? out_tile_w : 0;
          const bool two = (c + 32 < out_tile_w) && (ocol0 + 32 < ncols_out);
          const int wcols = two ? 64 : 32;
          if (ocol0 + wcols > ncols_out || acol0 + gofs + wcols > p.N) { c_done = c; break; }   // ragged: generic path
          c_done = c + 128;
          uint32_t raw[2][32];
          tmem_ld_32x32(taddr + c, raw[0]);
          if (two) tmem_ld_32x32(taddr + c + 32, raw[1]);
          if (resrow && !have_res) {     // first round of the tile (or after a ragged one): load now
            residual_prefetch(rr[0], resrow + ocol0);
            if (two) residual_prefetch(rr[1], resrow + ocol0 + 32);
          }
          have_res = false;
          if (res_tma) {
            if (!res_pending) res_issue(c);
            mbar_wait(my_res_bar, res_ph);
            res_ph ^= 1;
            res_pending = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              rr[0].u[q] = *reinterpret_cast<const uint4*>(res_buf + swz[q]);
              rr[1].u[q] = *reinterpret_cast<const uint4*>(res_buf + 2048 + swz[q]);
            }
            // The buffer may only be handed back to TMA once the eight loads above have RETURNED, not merely issued:
            // they queue behind this warp's outstanding global loads (bias / per-sample vectors), and a refill issued
            // right after them (the residual is usually L2 resident) overtook them - the "sparse wrong values" defect
            // of round 1: 16 B pieces of a round's residual read as the NEXT round's (zeros where that box lies
            // beyond N), gpurun_out/r02_s23 probe. The xor chain reads one word of every load (scoreboard wait).
            {
              // (ptxas deletes a dead xor chain even inside asm volatile: the result is stored to a scratch word)
              const uint32_t sink = rr[0].u[0].x ^ rr[0].u[1].x ^ rr[0].u[2].x ^ rr[0].u[3].x ^ rr[1].u[0].x ^
                                    rr[1].u[1].x ^ rr[1].u[2].x ^ rr[1].u[3].x;
              asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(res_sink)), "r"(sink) : "memory");
            }
            __syncwarp();                                   // every lane HAS the buffer's contents: it may be refilled
            if (lean_round(c + 128)) res_issue(c + 128);    // next round of this tile, behind this round's stores
          }
          tmem_ld_wait();
          float v[2][32];   // aliases raw: the accumulator registers are converted in place
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh == 1 && !two) break;
            if (EPF(ln_sums, 1, nullptr)) {   // (alpha == 1, no row bias: checked on the host)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh][j] = __uint_as_float(raw[hh][j]) * ln_a;
              fma_f32x32(p.ln_u + acol0 + hh * 32, ln_b, v[hh]);
            } else if (EPF(alpha, 2, 1.f) == 1.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh][j] = __uint_as_float(raw[hh][j]) + bm;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh][j] = fmaf(__uint_as_float(raw[hh][j]), EPF(alpha, 2, 1.f), bm);
            }
            if (p.bias) add_f32x32(p.bias + acol0 + hh * 32, v[hh]);
            if (rbb) add_f32x32(rbb + acol0 + hh * 32, v[hh]);
          }
          if (geglu) {
            uint32_t graw[2][32];
            tmem_ld_32x32(taddr + out_tile_w + c, graw[0]);
            if (two) tmem_ld_32x32(taddr + out_tile_w + c + 32, graw[1]);
            const int gcol0 = acol0 + out_tile_w;
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (hh == 1 && !two) break;
              float g[32];   // aliases graw[hh]
#pragma unroll
              for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(graw[hh][j]) * (EPF(ln_sums, 1, nullptr) ? ln_a : EPF(alpha, 2, 1.f));
              if (EPF(ln_sums, 1, nullptr)) fma_f32x32(p.ln_u + gcol0 + hh * 32, ln_b, g);
              if (p.bias) add_f32x32(p.bias + gcol0 + hh * 32, g);
#pragma unroll
              for (int j = 0; j < 32; j += 2) geglu_pair(v[hh][j], v[hh][j + 1], g[j], g[j + 1]);
            }
          } else if ((kActRt ? p.act : EPF(act, 1, (int)(kGeglu ? kActGeglu : kActNone))) == kActGeluTanh) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[0][j] = gelu_tanh_f(v[0][j]); v[1][j] = gelu_tanh_f(v[1][j]); }
          } else if ((kActRt ? p.act : EPF(act, 1, (int)(kGeglu ? kActGeglu : kActNone))) == kActSilu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[0][j] = silu_f(v[0][j]); v[1][j] = silu_f(v[1][j]); }
          }
          if (EPF(cap_pre, 2, nullptr))
            store_round<true>(&maps.cap_pre, ocol0, nullptr, 0, two, stg_warp, toggle, lane, swz, v[0], v[1], sc);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (hh == 1 && !two) break;
            if (csrow) mul_f32x32(csrow + ocol0 + hh * 32, v[hh]);
            if (resrow || res_tma) residual_add(rr[hh], v[hh]);
            if (EPF(out_scale, 1, 1.f) != 1.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[hh][j] *= EPF(out_scale, 1, 1.f);
            }
            if (EPF(row_sums, 1, nullptr)) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { rs_sum += v[hh][j]; rs_sq = fmaf(v[hh][j], v[hh][j], rs_sq); }
            }
            if (p.gn_sums) {
              const float part = gn_reduce(p, v[hh], row_ok, lane);
              if (c < 128) gn_run0[hh] += part; else gn_run1[hh] += part;
            }
          }
          if (resrow) {   // the residual registers are free again: fetch the next full round while this one is stored
            const int nc = c + 128, nocol0 = ocol0 + 128;
            if (nc + 64 <= out_tile_w && nocol0 + 64 <= ncols_out) {
              residual_prefetch(rr[0], resrow + nocol0);
              residual_prefetch(rr[1], resrow + nocol0 + 32);
              have_res = true;
            }
          }
          // destinations of the final value; a dtype / capture-segment boundary in the middle of the round
          // (all boundaries are multiples of 32 columns) splits it into two single-half rounds
          auto emit = [&](int col, bool tw, const float* va, const float* vb) {
            if (p.out && col >= EPF(out_f16_from, 2, (1 << 30))) {
              store_round<true>(&maps.out, col, nullptr, 0, tw, stg_warp, toggle, lane, swz, va, vb, sc);
              if (EPF(out2, 2, nullptr)) store_round<false>(&maps.out2, col, nullptr, 0, tw, stg_warp, toggle, lane, swz, va, vb, sc);
            } else if (p.out) {
              store_round<false>(&maps.out, col, EPF(out2, 2, nullptr) ? &maps.out2 : nullptr, col, tw, stg_warp, toggle, lane, swz, va,
                                 vb, sc);
            } else if (EPF(out2, 2, nullptr)) {
              store_round<false>(&maps.out2, col, nullptr, 0, tw, stg_warp, toggle, lane, swz, va, vb, sc);
            }
            if (EPF(num_cap, 2, 0) > 0) {
              const CUtensorMap* m0 = nullptr;
              const CUtensorMap* m1 = nullptr;
              int c0 = 0, c1 = 0;
#pragma unroll
              for (int s = 0; s < 3; ++s) {
                if (s < EPF(num_cap, 2, 0) && p.cap[s].ptr && col >= p.cap[s].col_begin && col < p.cap[s].col_end) {
                  if (!m0) { m0 = &maps.cap[s]; c0 = col - p.cap[s].col_begin; }
                  else if (!m1) { m1 = &maps.cap[s]; c1 = col - p.cap[s].col_begin; }
                }
              }
              if (m0) store_round<true>(m0, c0, m1, c1, tw, stg_warp, toggle, lane, swz, va, vb, sc);
            }
          };
          bool split = false;
          if (two) {
            const int mid = ocol0 + 32;
            split = (p.out && EPF(out_f16_from, 2, (1 << 30)) == mid);
#pragma unroll
            for (int s = 0; s < 3; ++s)
              if (s < EPF(num_cap, 2, 0) && p.cap[s].ptr && (p.cap[s].col_begin == mid || p.cap[s].col_end == mid)) split = true;
          }
          if (!split) emit(ocol0, two, v[0], v[1]);
          else {
            emit(ocol0, false, v[0], v[0]);
            emit(ocol0 + 32, false, v[1], v[1]);
          }
          if (EPF(out_f32, 1, nullptr) && row_ok) {
            direct_store32(p, v[0], v[0], row, ocol0, ocol0 + 32, out_batch_off, true);
            if (two) direct_store32(p, v[1], v[1], row, ocol0 + 32, ocol0 + 64, out_batch_off, true);
          }
        }
      }
      // ---- generic path: whatever the lean path left (ragged tiles, unaligned operands, no TMA store), 32 columns
      // at a time; this warp owns the 64-column groups with (group index & 1) == cset
      if (kLevel == 0)
      for (int c = c_done; c < out_tile_w; c += ((c & 32) ? 96 : 32)) {
        const int ocol0 = tc.n_tile * out_tile_w + c;
        if (ocol0 >= ncols_out) break;
        const int lim = min(ncols_out, ocol0 + 32);
        float v[32];
        load_activate32(p, taddr, c, out_tile_w, tc.n_tile, bm, row_ok, bidx, ln_a, ln_b, v);
        if (EPF(tma_store, 1, 1)) {
          // registers -> swizzled smem -> one TMA bulk store per destination (rows / columns clipped by the map)
          if (EPF(cap_pre, 2, nullptr)) stage_and_store<true>(&maps.cap_pre, stg_warp, toggle, lane, v, ocol0, sc);
          gate_residual32(p, v, row, bidx, ocol0, ncols_out, lim, row_ok);
          if (EPF(row_sums, 1, nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ocol0 + j < lim) { rs_sum += v[j]; rs_sq = fmaf(v[j], v[j], rs_sq); }
          }
          if (p.out) {
            if (ocol0 >= EPF(out_f16_from, 2, (1 << 30))) stage_and_store<true>(&maps.out, stg_warp, toggle, lane, v, ocol0, sc);
            else stage_and_store<false>(&maps.out, stg_warp, toggle, lane, v, ocol0, sc);
          }
          if (EPF(out2, 2, nullptr)) stage_and_store<false>(&maps.out2, stg_warp, toggle, lane, v, ocol0, sc);
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            if (s < EPF(num_cap, 2, 0) && p.cap[s].ptr && ocol0 >= p.cap[s].col_begin && ocol0 < p.cap[s].col_end)
              stage_and_store<true>(&maps.cap[s], stg_warp, toggle, lane, v, ocol0 - p.cap[s].col_begin, sc);
          }
          if (EPF(out_f32, 1, nullptr) && row_ok) direct_store32(p, v, v, row, ocol0, lim, out_batch_off, true);
        } else {
          float vpre[32];
          if (EPF(cap_pre, 2, nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vpre[j] = v[j];
          }
          gate_residual32(p, v, row, bidx, ocol0, ncols_out, lim, row_ok);
          if (EPF(row_sums, 1, nullptr)) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (ocol0 + j < lim) { rs_sum += v[j]; rs_sq = fmaf(v[j], v[j], rs_sq); }
          }
          if (row_ok) direct_store32(p, vpre, v, row, ocol0, lim, out_batch_off, false);
        }
      }
      if (EPF(row_sums, 1, nullptr) && row_ok) {   // this warp's column share of the row; the other column set / n-tiles add theirs
        atomicAdd(EPF(row_sums, 1, nullptr) + 2 * row, rs_sum);
        atomicAdd(EPF(row_sums, 1, nullptr) + 2 * row + 1, rs_sq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && cta_rank != 0) mbar_arrive_remote(&tempty_bar[as], 0);   // the leader's MMA issuer waits for both
        else mbar_arrive(&tempty_bar[as]);
      }
      if (++as == acc_stages) { as = 0; aph ^= 1; }
    }
    if (p.gn_sums && gn_key >= 0) gn_flush_run();
    __syncwarp();
    if (elect_one()) bulk_wait<0>();   // all bulk stores of this warp have completed before the CTA retires
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // pair: neither CTA retires while its peer still uses it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host side
static int g_num_sms = 0;
static int g_pair_ctas = 0;   // CTAs the device keeps resident as clusters of 2 (<= g_num_sms)

static cudaError_t gemm_init_once() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(gemm_tcgen05_kernel<2, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes);
  if (e != cudaSuccess) return e;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g_num_sms & ~1);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmemBytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, gemm_tcgen05_kernel<2, false>, &cfg) == cudaSuccess && nclusters > 0)
    g_pair_ctas = 2 * nclusters;
  else
    g_pair_ctas = 0;
  (void)cudaGetLastError();
  if (g_pair_ctas > g_num_sms) g_pair_ctas = g_num_sms & ~1;
  done = true;
  return cudaSuccess;
}

cudaError_t launch_gemm(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream) {
  cudaError_t e = gemm_init_once();
  if (e != cudaSuccess) return e;
  const int cg = p.cta_group == 2 ? 2 : 1;
  const int num_m_units = (p.num_m_tiles + cg - 1) / cg;
  const int total_units = p.batch * num_m_units * p.num_n_tiles;
  if (total_units <= 0) return cudaSuccess;
  if (cg == 1) {
    const int grid = total_units < g_num_sms ? total_units : g_num_sms;
    if (p.act == kActGeglu)
      return launch_pdl(gemm_tcgen05_kernel<1, true>, dim3(grid), dim3(kGemmThreads), (size_t)kGemmSmemBytes, stream, maps, p);
    return launch_pdl(gemm_tcgen05_kernel<1, false>, dim3(grid), dim3(kGemmThreads), (size_t)kGemmSmemBytes, stream, maps, p);
  }
  if (g_pair_ctas < 2) return cudaErrorInvalidConfiguration;
  const int grid = 2 * total_units < g_pair_ctas ? 2 * total_units : g_pair_ctas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  // specialised epilogues (see kLevel): GDF_EPI_LEVEL=0 forces the general kernel, =1 stops at "lean" (A/B timing)
  static const int max_level = getenv("GDF_EPI_LEVEL") ? atoi(getenv("GDF_EPI_LEVEL")) : 2;
  const bool geglu = p.act == kActGeglu;
  const int out_w = geglu ? p.block_n / 2 : p.block_n;
  const bool act_rt = p.act == kActGeluTanh || p.act == kActSilu;
  const bool lean = max_level >= 1 && p.sk_pieces <= 1 && (geglu || p.act == kActNone || act_rt) && !p.ln_sums && !p.row_sums && !p.bias_m &&
                    p.out_scale == 1.f && !p.out_f32 && p.fast_epi && p.tma_store &&
                    p.N % p.block_n == 0 && out_w % 32 == 0 && p.n_out == (geglu ? p.N / 2 : p.N);
  if (geglu) {
    if (lean) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, true, 1>, maps, p);
    return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, true>, maps, p);
  }
  if (lean && act_rt) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, false, 1, true>, maps, p);
  const bool simple = lean && max_level >= 2 && p.alpha == 1.f && !p.col_scale && !p.row_batch_bias && !p.out2 && !p.cap_pre &&
                      p.num_cap == 0 && p.out && p.out_f16_from >= (1 << 30);
  if (simple) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, false, 2>, maps, p);
  if (lean) return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, false, 1>, maps, p);
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, false>, maps, p);
}

int gemm_resident_groups(int cta_group) {
  if (gemm_init_once() != cudaSuccess) return 0;
  return cta_group == 2 ? g_pair_ctas / 2 : g_num_sms;
}

int gemm_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace gdf
