// Persistent warp-specialised tcgen05 GEMM / implicit-GEMM convolution kernel (see gemm_sm100.cuh).
//
// Roles (256 threads, 1 CTA per SM, TMEM 512 columns = 2 accumulator stages of 128 lanes x 256 fp32):
//   warp 0  : TMA producer  (A tile 128x64 bf16, B tile block_n x 64 bf16, 128B swizzle, 4-stage ring)
//   warp 1  : MMA issuer    (one elected lane issues tcgen05.mma 128 x block_n x 16, commits to mbarriers)
//   warp 2  : TMEM allocator / deallocator
//   warps 4-7: epilogue     (tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global)
// Pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), static persistent tile loop.
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace gdf {

struct TileCoord {
  int m_tile, n_tile, bz;
};
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

// Fused epilogue over one chunk of 32 output columns of one row.
//   v[]    : activated accumulator values (alpha, biases and activation already applied)
//   ocol0  : first output column of the chunk, ncols_out: output width
__device__ __forceinline__ void epilogue_store_chunk(const GemmParams& p, float (&v)[32], long long row, int bidx,
                                                     int ocol0, int ncols_total, int ncols_out,
                                                     long long out_batch_off) {
  // ncols_total: full output width (indexing of per-sample vectors); ncols_out: exclusive column limit of this chunk
  const bool full = (ocol0 + 32 <= ncols_out);
  // ---- capture before residual ("increment")
  if (p.cap_pre) {
    __half* dst = p.cap_pre + row * p.ld_cap_pre + ocol0;
    if (full && (p.ld_cap_pre % 8 == 0)) {
      store16_f16(dst, v);
      store16_f16(dst + 16, v + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < ncols_out) dst[j] = __float2half_rn(v[j]);
    }
  }
  // ---- per-sample column gate
  if (p.col_scale) {
    const float* g = p.col_scale + (long long)bidx * ncols_total + ocol0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (ocol0 + j < ncols_out) v[j] *= __ldg(g + j);
  }
  // ---- residual
  if (p.residual) {
    const __nv_bfloat16* r = p.residual + row * p.ld_res + ocol0;
    if (full && (p.ld_res % 8 == 0)) {
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
        if (ocol0 + j < ncols_out) v[j] += __bfloat162float(r[j]);
    }
  }
  if (p.out_scale != 1.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
  }
  // ---- destinations
  if (p.out) {
    __nv_bfloat16* dst = p.out + out_batch_off + row * p.ld_out + ocol0;
    if (full && (p.ld_out % 8 == 0)) {
      store16_bf16(dst, v);
      store16_bf16(dst + 16, v + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < ncols_out) dst[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (p.out2) {
    __nv_bfloat16* dst = p.out2 + row * p.ld_out2 + ocol0;
    if (full && (p.ld_out2 % 8 == 0)) {
      store16_bf16(dst, v);
      store16_bf16(dst + 16, v + 16);
    } else {
      for (int j = 0; j < 32; ++j)
        if (ocol0 + j < ncols_out) dst[j] = __float2bfloat16_rn(v[j]);
    }
  }
  if (p.out_f32) {
    float* dst = p.out_f32 + row * p.ld_out_f32 + ocol0;
    for (int j = 0; j < 32; ++j)
      if (ocol0 + j < ncols_out) dst[j] = v[j];
  }
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

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kStageBytesA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kStageBytesA + kStageBytesB));
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [kAccStages]
  uint64_t* tempty_bar = bars + 2 * kStages + kAccStages;  // [kAccStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.batch * p.num_m_tiles * p.num_n_tiles;
  const uint32_t stage_tx_bytes = (kBlockM + p.block_n) * kBlockK * 2;

  if (warp == 0) {
    // ===================================================== TMA producer
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      int x0 = 0, y0 = 0, b0 = 0;
      if (p.a_mode != kALinear) {
        const int xt = tc.m_tile % p.tiles_x;
        const int r = tc.m_tile / p.tiles_x;
        const int yt = r % p.tiles_y;
        const int bt = r / p.tiles_y;
        x0 = xt * p.tw;
        y0 = yt * p.th;
        b0 = bt * p.tb;
      }
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(&full_bar[s], stage_tx_bytes);
          uint8_t* a_dst = sA + s * kStageBytesA;
          uint8_t* b_dst = sB + s * kStageBytesB;
          if (p.a_mode == kALinear) {
            tma_load_3d(a_dst, &map_a, &full_bar[s], kb * kBlockK, tc.m_tile * kBlockM, p.a_batched ? tc.bz : 0);
          } else {
            const int tap = kb / p.cin_blocks;
            const int cb = kb - tap * p.cin_blocks;
            const int ky = tap / 3, kx = tap - 3 * ky;
            if (p.a_mode == kAConvS1) {
              tma_load_4d(a_dst, &map_a, &full_bar[s], cb * kBlockK, x0 + kx - 1, y0 + ky - 1, b0);
            } else {
              // stride-2: input viewed as (B, H, 2, W, 2*Cin) with H, W the OUTPUT extents;
              // input row 2*oy + ky - pad_lo -> parity (t & 1), half-row oy + (t >> 1)
              const int ty = ky - p.pad_lo, tx = kx - p.pad_lo;
              const int ypar = ty & 1, yoff = ty >> 1;
              const int xpar = tx & 1, xoff = tx >> 1;
              tma_load_5d(a_dst, &map_a, &full_bar[s], xpar * p.cin_blocks * kBlockK + cb * kBlockK, x0 + xoff, ypar,
                          y0 + yoff, b0);
            }
          }
          tma_load_3d(b_dst, &map_b, &full_bar[s], kb * kBlockK, tc.n_tile * p.block_n, p.b_batched ? tc.bz : 0);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    const uint32_t idesc = umma_idesc_bf16(kBlockM, p.block_n);
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&tempty_bar[as], aph ^ 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * kMaxBlockN;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + s * kStageBytesA));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + s * kStageBytesB));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_f16_ss(tmem_acc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tfull_bar[as]);
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      if (++as == kAccStages) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue
    const int ew = warp - 4;  // == warp % 4: TMEM lane quadrant this warp may access
    const int r_in_tile = ew * 32 + lane;
    int as = 0;
    uint32_t aph = 0;
    const bool geglu = (p.act == kActGeglu);
    const int ncols_out = p.n_out;
    const int out_tile_w = geglu ? p.block_n / 2 : p.block_n;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      // ---- row of this thread
      long long row;
      bool row_ok;
      if (p.a_mode == kALinear) {
        row = (long long)tc.m_tile * kBlockM + r_in_tile;
        row_ok = row < p.M;
      } else {
        const int xt = tc.m_tile % p.tiles_x;
        const int r = tc.m_tile / p.tiles_x;
        const int yt = r % p.tiles_y;
        const int bt = r / p.tiles_y;
        const int tx = r_in_tile % p.tw;
        const int r2 = r_in_tile / p.tw;
        const int ty = r2 % p.th;
        const int tbi = r2 / p.th;
        const int b = bt * p.tb + tbi;
        row = ((long long)b * p.H + (yt * p.th + ty)) * p.W + (xt * p.tw + tx);
        row_ok = b < p.B_img;
      }
      const int bidx = (p.rows_per_batch > 0) ? (int)(row / p.rows_per_batch) : 0;
      const long long out_batch_off = (long long)tc.bz * p.out_batch_stride;
      const float bm = (p.bias_m && row_ok) ? __ldg(p.bias_m + row) : 0.f;

      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(ew * 32) << 16) + as * kMaxBlockN;

      for (int c = 0; c < out_tile_w; c += 32) {
        uint32_t raw[32];
        float v[32];
        tmem_ld_32x32(taddr + c, raw);
        tmem_ld_wait();
        const int acol0 = tc.n_tile * p.block_n + c;       // accumulator column (bias index)
        const int ocol0 = tc.n_tile * out_tile_w + c;      // output column
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(raw[j]) * p.alpha + bm;
          if (p.bias && acol0 + j < p.N) x += __ldg(p.bias + acol0 + j);
          v[j] = x;
        }
        if (p.row_batch_bias && row_ok) {
          const float* rb = p.row_batch_bias + (long long)bidx * p.N + acol0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (acol0 + j < p.N) v[j] += __ldg(rb + j);
        }
        if (geglu) {
          uint32_t graw[32];
          tmem_ld_32x32(taddr + out_tile_w + c, graw);
          tmem_ld_wait();
          const int gcol0 = acol0 + out_tile_w;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float g = __uint_as_float(graw[j]) * p.alpha;
            if (p.bias && gcol0 + j < p.N) g += __ldg(p.bias + gcol0 + j);
            v[j] *= gelu_erf_f(g);
          }
        } else if (p.act == kActGeluTanh) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
        } else if (p.act == kActSilu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
        }
        if (row_ok && ocol0 < ncols_out) {
          const int lim = min(ncols_out, ocol0 + min(32, out_tile_w - c));
          epilogue_store_chunk(p, v, row, bidx, ocol0, ncols_out, lim, out_batch_off);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == kAccStages) { as = 0; aph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------ host side
static int g_num_sms = 0;

cudaError_t launch_gemm(const CUtensorMap& map_a, const CUtensorMap& map_b, const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kGemmSmemBytes);
    if (e != cudaSuccess) return e;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    attr_set = true;
  }
  const int total_tiles = p.batch * p.num_m_tiles * p.num_n_tiles;
  if (total_tiles <= 0) return cudaSuccess;
  const int grid = total_tiles < g_num_sms ? total_tiles : g_num_sms;
  gemm_tcgen05_kernel<<<grid, kGemmThreads, kGemmSmemBytes, stream>>>(map_a, map_b, p);
  return cudaGetLastError();
}

}  // namespace gdf
