// First convolution of the VAE encoder (3 -> N channels, 3x3, pad 1) fused from the fp32 NCHW image to the bf16 NHWC
// activation: [diffusers AutoencoderKL Encoder.conv_in], called from pipeline_pixart_sigma.py:644 (vae.encode).
//
// A K = 27 contraction is far too thin for the general implicit-GEMM kernel (it needs 64-channel TMA boxes), so round 1
// materialised a [M, 64] im2col operand (1.07 GB written + read again per 8-image batch) and ran a K = 64 GEMM:
// 0.83 + 1.70 ms per step against 0.34 ms of unavoidable traffic (image read once, activation written once). Here the
// operand tile never leaves the SM:
//   warps 0-3  : builders - thread p gathers the 27 taps of pixel p of a 128-pixel row segment straight from the image
//                (coalesced along x, reuse through L1), rounds to bf16 and writes row p of the A tile into shared
//                memory in the canonical K-major SWIZZLE_128B layout (K padded to 32), 3-stage ring
//   warp 4     : MMA issuer - two tcgen05.mma (128 x N x 16) per tile into one of two TMEM accumulator stages
//   warps 5-8  : epilogue - tcgen05.ld, + bias, GroupNorm statistics of the output for the GroupNorm that consumes it
//                (running per-thread sums over the tiles of one image, one shuffle reduction + atomics per image),
//                bf16 pack into a SWIZZLE_128B staging tile, one TMA bulk store per 64 channels (double buffered)
// Persistent: one CTA per SM walks tiles bid, bid + grid, ...; the weights (N x 32, 8 KB) are loaded once per CTA.
// HBM bound: algorithmic bytes = 12 B (image) + 2 N B (activation) per pixel.
#include <stdlib.h>
#include "ops.h"

namespace gdf {

constexpr int kCiThreads = 288;          // 9 warps
constexpr int kCiStages = 3;
constexpr int kCiATile = 128 * 128;      // 128 rows x 128 B (64 bf16 slots per row, 32 used)
constexpr int kCiOffB = kCiStages * kCiATile;             // weights: N (<= 128) rows x 128 B
constexpr int kCiOffStg = kCiOffB + 128 * 128;            // output staging: 2 buffers x 2 halves x (128 rows x 128 B)
constexpr int kCiOffBias = kCiOffStg + 2 * 2 * kCiATile;  // fp32 [128]
constexpr int kCiOffBar = kCiOffBias + 512;
constexpr int kCiSmem = kCiOffBar + 256 + 1024;

struct ConvInParams {
  const float* img;        // (B, 3, H, W) fp32
  const float* bias;       // [N] fp32
  int B, H, W, N;          // W % 128 == 0, N in {64, 128}
  int num_tiles;           // B * H * (W / 128)
  float* gn_sums;          // optional fp32 [B][G][2]; groups of cpg channels
  int gn_cpg_log2, gn_groups;
};

// kCpgLog2: log2 of the channels per GroupNorm group (2 / 3 / 4), compile time so that the statistics stay in registers
// (with a run-time group width the per-group loops indexed the value array dynamically: 160 B of stack, 470 LDL / STL
// and ~2800 instructions per tile and epilogue thread - the first version ran at 6.7 us per tile, 2.98 ms per step).
template <int kCpgLog2>
__global__ void __launch_bounds__(kCiThreads, 1)
conv_in_tcgen05_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                       const ConvInParams p) {
  extern __shared__ uint8_t ci_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ci_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCiOffBar);
  uint64_t* a_full = bars;                       // [stages] count 4 (one arrive per builder warp)
  uint64_t* a_empty = a_full + kCiStages;        // [stages] count 1 (tcgen05.commit)
  uint64_t* acc_full = a_empty + kCiStages;      // [2] count 1 (tcgen05.commit)
  uint64_t* acc_empty = acc_full + 2;            // [2] count 4 (one arrive per epilogue warp)
  uint64_t* w_full = acc_empty + 2;              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* bias_s = reinterpret_cast<float*>(smem + kCiOffBias);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N;
  const int tiles_x = p.W >> 7;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_out);
    for (int i = 0; i < kCiStages; ++i) {
      mbar_init(&a_full[i], 4);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (threadIdx.x < 128) bias_s[threadIdx.x] = (threadIdx.x < N) ? p.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp < 4) {
    // ================================================= builders: image -> A tile (im2col rows in shared memory)
    const int px = threadIdx.x;                  // pixel of the tile = row of the A tile
    const uint32_t swz = (px & 7) << 4;
    int s = 0;
    uint32_t ph = 0;
    // k = (ky*3 + kx)*3 + c, 27 values (+ 5 zeros) -> 16 packed bf16 pairs. The gather of tile i + 1 is issued before
    // tile i is written: two tiles of loads in flight per thread (the loop was paced by one DRAM round trip per tile).
    auto gather = [&](int tile, float* v) {
      const int xt = tile % tiles_x;
      const int r = tile / tiles_x;
      const int y = r % p.H, b = r / p.H;
      const int x = xt * 128 + px;
      const float* ib = p.img + (long long)b * 3 * p.H * p.W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        const bool yok = yy >= 0 && yy < p.H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          const bool ok = yok && xx >= 0 && xx < p.W;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            v[(ky * 3 + kx) * 3 + c] = ok ? __ldg(ib + ((long long)c * p.H + yy) * p.W + xx) : 0.f;
        }
      }
    };
    float v[32], vn[27];
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.f;
    if (blockIdx.x < p.num_tiles) gather(blockIdx.x, v);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int next = tile + gridDim.x;
      if (next < p.num_tiles) gather(next, vn);
      mbar_wait(&a_empty[s], ph ^ 1);
      const uint32_t row = smem_u32(smem + s * kCiATile) + px * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        st_shared_v4(row + ((ch << 4) ^ swz), pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]),
                     pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]), pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]),
                     pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]));
      fence_proxy_async_smem();      // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      if (++s == kCiStages) { s = 0; ph ^= 1; }
#pragma unroll
      for (int i = 0; i < 27; ++i) v[i] = vn[i];
    }
  } else if (warp == 4) {
    // ================================================= MMA issuer
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, N * 128);
      tma_load_2d(smem + kCiOffB, &map_w, w_full, 0, 0);
    }
    mbar_wait(w_full, 0);
    const uint32_t idesc = umma_idesc_bf16(128, N, 0);
    const uint64_t db = umma_desc_kmajor_sw128(smem_u32(smem + kCiOffB));
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[as], aph ^ 1);
      mbar_wait(&a_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem + s * kCiATile));
        umma_f16_ss(tmem_base + as * 128, da, db, idesc, 0);
        umma_f16_ss(tmem_base + as * 128, da + 2, db + 2, idesc, 1);
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[as]);
      }
      __syncwarp();
      if (++s == kCiStages) { s = 0; ph ^= 1; }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else {
    // ================================================= epilogue (warps 5-8: TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_off = uint32_t(quad * 32) << 16;
    const uint32_t swz = (r_in_tile & 7) << 4;
    const int nch = N >> 5;                       // 32-column chunks (2 or 4)
    constexpr int kCpg = 1 << kCpgLog2;
    constexpr int kGpc = 32 >> kCpgLog2;          // groups per 32-column chunk (8 / 4 / 2)
    const bool gn = p.gn_sums != nullptr;
    // running GroupNorm sums of this thread's row over the tiles of one image: [32-col chunk][group in chunk][sum, sq]
    float gacc[4][kGpc][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int g2 = 0; g2 < kGpc; ++g2) gacc[a][g2][0] = gacc[a][g2][1] = 0.f;
    int gn_img = -1;
    auto gn_flush = [&]() {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a >= nch) continue;
#pragma unroll
        for (int g2 = 0; g2 < kGpc; ++g2) {
          float sm = gacc[a][g2][0], sq = gacc[a][g2][1];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          if (lane == 0) {
            float* dst = p.gn_sums + ((long long)gn_img * p.gn_groups + a * kGpc + g2) * 2;
            atomicAdd(dst, sm);
            atomicAdd(dst + 1, sq);
          }
          gacc[a][g2][0] = gacc[a][g2][1] = 0.f;
        }
      }
    };
    int as = 0, sb = 0;
    uint32_t aph = 0;
    const int tiles_per_img = tiles_x * p.H;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img;
      if (gn && img != gn_img) {
        if (gn_img >= 0) gn_flush();
        gn_img = img;
      }
      mbar_wait(&acc_full[as], aph);
      tc_fence_after();
      // the staging buffer `sb` was handed to TMA two tiles ago: its reads must have completed
      if (warp == 5 && lane == 0) bulk_wait_read<1>();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* stg = smem + kCiOffStg + sb * 2 * kCiATile;
#pragma unroll
      for (int a2 = 0; a2 < 4; a2 += 2) {          // two 32-column chunks per round: both TMEM loads in flight
        if (a2 >= nch) continue;
        uint32_t raw[2][32];
        tmem_ld_32x32(tmem_base + lane_off + as * 128 + a2 * 32, raw[0]);
        tmem_ld_32x32(tmem_base + lane_off + as * 128 + a2 * 32 + 32, raw[1]);
        tmem_ld_wait();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int a = a2 + hh;
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(bias_s + a * 32 + q * 4);   // broadcast LDS.128
            v[q * 4 + 0] = __uint_as_float(raw[hh][q * 4 + 0]) + bq.x;
            v[q * 4 + 1] = __uint_as_float(raw[hh][q * 4 + 1]) + bq.y;
            v[q * 4 + 2] = __uint_as_float(raw[hh][q * 4 + 2]) + bq.z;
            v[q * 4 + 3] = __uint_as_float(raw[hh][q * 4 + 3]) + bq.w;
          }
          if (gn) {
#pragma unroll
            for (int g2 = 0; g2 < kGpc; ++g2) {
              float sm = 0.f, sq = 0.f;
#pragma unroll
              for (int j = 0; j < kCpg; ++j) {
                sm += v[g2 * kCpg + j];
                sq = fmaf(v[g2 * kCpg + j], v[g2 * kCpg + j], sq);
              }
              gacc[a][g2][0] += sm;
              gacc[a][g2][1] += sq;
            }
          }
          // 32 channels = 64 B = 4 chunks of 16 B in the 128 B row of half (a >> 1), chunk index (a & 1) * 4 + q
          const uint32_t row = smem_u32(stg + (a >> 1) * kCiATile) + r_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_shared_v4(row + ((((a & 1) * 4 + q) << 4) ^ swz), pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]),
                         pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]), pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]),
                         pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        }
      }
      // accumulator stage drained: the issuer may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (warp == 5 && lane == 0) {
        const int row0 = tile * 128;             // tiles are 128 consecutive rows of the [M, N] output
        tma_store_2d(&map_out, stg, 0, row0);
        if (N > 64) tma_store_2d(&map_out, stg + kCiATile, 64, row0);
        bulk_commit();
      }
      sb ^= 1;
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (gn && gn_img >= 0) gn_flush();
    if (warp == 5 && lane == 0) bulk_wait<0>();   // every store has landed before the CTA (and its smem) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// img: fp32 NCHW (B, 3, H, W); w_packed: bf16 [N_pad][64] with k = (ky*3+kx)*3 + c (launch_pack_conv_weight, k_pad 64);
// out: bf16 [B*H*W, N]. Returns false (nothing launched) when the shape is outside what the kernel serves.
bool conv_in_fused_supported(int Cin, int N, int W) { return Cin == 3 && (N == 64 || N == 128) && W % 128 == 0; }

int launch_conv_in_fused(const float* img, const bf16* w_packed, const float* bias, bf16* out, int B, int H, int W, int N,
                         float* gn_sums, int gn_cpg, int gn_groups, cudaStream_t stream) {
  if (!conv_in_fused_supported(3, N, W)) return fail(GDF_ERR_UNSUPPORTED, "conv_in_fused: N=%d W=%d", N, W);
  if (gn_sums && !(gn_cpg == 4 || gn_cpg == 8 || gn_cpg == 16))
    return fail(GDF_ERR_UNSUPPORTED, "conv_in_fused: GroupNorm groups of %d channels", gn_cpg);
  static bool attr = false;
  if (!attr) {
    GDF_CUDA(cudaFuncSetAttribute(conv_in_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCiSmem));
    GDF_CUDA(cudaFuncSetAttribute(conv_in_tcgen05_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCiSmem));
    GDF_CUDA(cudaFuncSetAttribute(conv_in_tcgen05_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCiSmem));
    attr = true;
  }
  CUtensorMap mw, mo;
  {
    uint64_t dims[2] = {64, (uint64_t)N};
    uint64_t str[1] = {64 * 2};
    uint32_t box[2] = {64, (uint32_t)N};
    GDF_TRY(make_tmap_bf16(&mw, w_packed, 2, dims, str, box));
  }
  {
    const uint64_t M = (uint64_t)B * H * W;
    uint64_t dims[2] = {(uint64_t)N, M};
    uint64_t str[1] = {(uint64_t)N * 2};
    uint32_t box[2] = {64, 128};
    GDF_TRY(make_tmap_bf16(&mo, out, 2, dims, str, box));
  }
  ConvInParams p;
  p.img = img;
  p.bias = bias;
  p.B = B;
  p.H = H;
  p.W = W;
  p.N = N;
  p.num_tiles = B * H * (W / 128);
  p.gn_sums = gn_sums;
  p.gn_cpg_log2 = gn_cpg == 16 ? 4 : gn_cpg == 8 ? 3 : 2;
  p.gn_groups = gn_groups;
  const int sms = gemm_num_sms();
  dim3 grid(p.num_tiles < sms ? p.num_tiles : sms);
  if (p.gn_cpg_log2 == 2)
    GDF_CUDA(launch_pdl(conv_in_tcgen05_kernel<2>, grid, dim3(kCiThreads), (size_t)kCiSmem, stream, mw, mo, p));
  else if (p.gn_cpg_log2 == 3)
    GDF_CUDA(launch_pdl(conv_in_tcgen05_kernel<3>, grid, dim3(kCiThreads), (size_t)kCiSmem, stream, mw, mo, p));
  else
    GDF_CUDA(launch_pdl(conv_in_tcgen05_kernel<4>, grid, dim3(kCiThreads), (size_t)kCiSmem, stream, mw, mo, p));
  return GDF_OK;
}

}  // namespace gdf
