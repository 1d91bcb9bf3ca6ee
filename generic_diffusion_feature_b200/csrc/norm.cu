// Fused GroupNorm(+SiLU) and LayerNorm(+AdaLN modulation) for NHWC / token-major bf16 activations.
// HBM-bound: 128-bit coalesced accesses, fp32 statistics, warp-shuffle + shared-memory reductions.
// Reference arithmetic: torch.nn.GroupNorm + SiLU in feature/diffusers/models/resnet.py:327-328,351,363,
// unet/unet_2d_condition.py:1305-1306, transformers/transformer_2d.py:484; nn.LayerNorm and the
// ada_norm_single modulation in feature/diffusers/models/attention.py:498-503,539,565,570-572.
#include <stdlib.h>
#include "ops.h"

namespace gdf {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 256;
constexpr int kGnMaxSlots = 2;  // channel slots (8 channels each) per thread: supports C <= 4096

struct GnMap {      // thread -> (channel slot(s), pixel lane) mapping shared by both GroupNorm kernels
  int c8;           // C / 8
  int lanes;        // pixel lanes per block
  int slots;        // slots per thread
};
__host__ __device__ inline GnMap gn_map(int C) {
  GnMap m;
  m.c8 = C / 8;
  if (m.c8 <= kGnThreads) {
    m.lanes = kGnThreads / m.c8;
    m.slots = 1;
  } else {
    m.lanes = 1;
    m.slots = (m.c8 + kGnThreads - 1) / kGnThreads;
  }
  return m;
}

bool deterministic_mode() {
  static const bool on = [] { const char* e = getenv("GDF_DETERMINISTIC"); return e && e[0] == '1'; }();
  return on;
}

size_t gn_workspace_floats(int B, int G) { return (size_t)B * kGnMaxChunks * G * 2 + (size_t)B * G * 2; }

// Pass 1: per (chunk, image) partial sum / sum of squares per group.
__global__ void __launch_bounds__(kGnThreads)
groupnorm_stats_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int HW, int C, int G, int nchunks) {
  __shared__ float s_sum[64], s_sq[64];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y, chunk = blockIdx.x;
  const GnMap m = gn_map(C);
  const int cpg = C / G;
  if (threadIdx.x < 64) {
    s_sum[threadIdx.x] = 0.f;
    s_sq[threadIdx.x] = 0.f;
  }
  __syncthreads();
  const int pix_per_chunk = (HW + nchunks - 1) / nchunks;
  const int p0 = chunk * pix_per_chunk;
  const int p1 = min(HW, p0 + pix_per_chunk);
  const int lane_p = (m.slots == 1) ? (threadIdx.x / m.c8) : 0;
  const bool active = (m.slots > 1) || (lane_p < m.lanes);
  for (int si = 0; si < m.slots; ++si) {
    const int slot = (m.slots == 1) ? (threadIdx.x % m.c8) : (threadIdx.x + si * kGnThreads);
    if (!active || slot >= m.c8) continue;
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
    const bf16* base = x + ((long long)b * HW) * C + slot * 8;
    auto acc = [&](const uint4& u) {
      float2 f;
      f = unpack_bf16x2(u.x); sum[0] += f.x; sq[0] += f.x * f.x; sum[1] += f.y; sq[1] += f.y * f.y;
      f = unpack_bf16x2(u.y); sum[2] += f.x; sq[2] += f.x * f.x; sum[3] += f.y; sq[3] += f.y * f.y;
      f = unpack_bf16x2(u.z); sum[4] += f.x; sq[4] += f.x * f.x; sum[5] += f.y; sq[5] += f.y * f.y;
      f = unpack_bf16x2(u.w); sum[6] += f.x; sq[6] += f.x * f.x; sum[7] += f.y; sq[7] += f.y * f.y;
    };
    int p = p0 + lane_p;
    for (; p + 3 * m.lanes < p1; p += 4 * m.lanes) {   // four independent 128-bit loads in flight per thread
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + (long long)p * C));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (long long)(p + m.lanes) * C));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (long long)(p + 2 * m.lanes) * C));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (long long)(p + 3 * m.lanes) * C));
      acc(u0); acc(u1); acc(u2); acc(u3);
    }
    for (; p < p1; p += m.lanes) acc(__ldg(reinterpret_cast<const uint4*>(base + (long long)p * C)));
    // fold the 8 channels into their groups (a slot may straddle group boundaries)
    int g_prev = (slot * 8) / cpg;
    float gs = 0.f, gq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (slot * 8 + j) / cpg;
      if (g != g_prev) {
        atomicAdd(&s_sum[g_prev], gs);
        atomicAdd(&s_sq[g_prev], gq);
        gs = gq = 0.f;
        g_prev = g;
      }
      gs += sum[j];
      gq += sq[j];
    }
    atomicAdd(&s_sum[g_prev], gs);
    atomicAdd(&s_sq[g_prev], gq);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    float* dst = partial + (((long long)b * nchunks + chunk) * G + threadIdx.x) * 2;
    dst[0] = s_sum[threadIdx.x];
    dst[1] = s_sq[threadIdx.x];
  }
}

// Deterministic variant of pass 1 (GDF_DETERMINISTIC=1): one block per (chunk, group, image), a fixed element -> thread
// assignment and a fixed-order tree reduction in shared memory instead of shared-memory atomics, so that two runs give
// bit-identical statistics (and therefore bit-identical features). Strided reads: slower, opt-in.
__global__ void __launch_bounds__(kGnThreads)
groupnorm_stats_det_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int HW, int C, int G, int nchunks) {
  __shared__ float s_a[kGnThreads], s_b[kGnThreads];
  pdl_wait();
  pdl_trigger();
  const int chunk = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int cpg = C / G;
  const int pix_per_chunk = (HW + nchunks - 1) / nchunks;
  const int p0 = chunk * pix_per_chunk;
  const int p1 = min(HW, p0 + pix_per_chunk);
  const long long n = (long long)(p1 > p0 ? p1 - p0 : 0) * cpg;
  const bf16* base = x + ((long long)b * HW) * C + g * cpg;
  float sm = 0.f, sq = 0.f;
  for (long long i = threadIdx.x; i < n; i += kGnThreads) {
    const long long p = p0 + i / cpg;
    const int c = (int)(i % cpg);
    const float v = __bfloat162float(base[p * C + c]);
    sm += v;
    sq = fmaf(v, v, sq);
  }
  s_a[threadIdx.x] = sm;
  s_b[threadIdx.x] = sq;
  __syncthreads();
  for (int o = kGnThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* dst = partial + (((long long)b * nchunks + chunk) * G + g) * 2;
    dst[0] = s_a[0];
    dst[1] = s_b[0];
  }
}

// Pass 1b: one block per image, one warp per group: lanes stride over the per-chunk partials, double-precision
// shuffle reduction -> (mean, rstd) per group.
__global__ void groupnorm_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats, int HW, int C,
                                          int G, int nchunks, float eps) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, lane = threadIdx.x & 31;
  for (int g = threadIdx.x >> 5; g < G; g += blockDim.x >> 5) {
    double s = 0.0, q = 0.0;
    const float2* src = reinterpret_cast<const float2*>(partial) + ((long long)b * nchunks * G + g);
    for (int c = lane; c < nchunks; c += 32) {
      const float2 v = __ldg(src + (long long)c * G);
      s += (double)v.x;
      q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double n = (double)HW * (C / G);
      const double mean = s / n;
      double var = q / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[(b * G + g) * 2] = (float)mean;
      stats[(b * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

// Pass 2: normalise, affine, optional SiLU.
__global__ void __launch_bounds__(kGnThreads)
groupnorm_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ partial, int HW, int C, int G,
                       int nchunks, int silu) {
  __shared__ float s_mean[64], s_rstd[64];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y, chunk = blockIdx.x;
  const GnMap m = gn_map(C);
  const int cpg = C / G;
  if (threadIdx.x < G) {
    s_mean[threadIdx.x] = partial[(b * G + threadIdx.x) * 2];       // `partial` is the finalised stats array here
    s_rstd[threadIdx.x] = partial[(b * G + threadIdx.x) * 2 + 1];
  }
  __syncthreads();
  const int pix_per_chunk = (HW + nchunks - 1) / nchunks;
  const int p0 = chunk * pix_per_chunk;
  const int p1 = min(HW, p0 + pix_per_chunk);
  const int lane_p = (m.slots == 1) ? (threadIdx.x / m.c8) : 0;
  const bool active = (m.slots > 1) || (lane_p < m.lanes);
  for (int si = 0; si < m.slots; ++si) {
    const int slot = (m.slots == 1) ? (threadIdx.x % m.c8) : (threadIdx.x + si * kGnThreads);
    if (!active || slot >= m.c8) continue;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = slot * 8 + j;
      const int g = c / cpg;
      const float a = s_rstd[g] * __ldg(gamma + c);
      sc[j] = a;
      sh[j] = __ldg(beta + c) - s_mean[g] * a;
    }
    const long long base = ((long long)b * HW) * C + slot * 8;
    auto emit = [&](const uint4& u, int p) {
      float v[8];
      float2 f;
      f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
      f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
      f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
      f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = fmaf(v[j], sc[j], sh[j]);
        v[j] = silu ? silu_f(t) : t;
      }
      uint4 o;
      o.x = pack_bf16x2(v[0], v[1]);
      o.y = pack_bf16x2(v[2], v[3]);
      o.z = pack_bf16x2(v[4], v[5]);
      o.w = pack_bf16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(y + base + (long long)p * C) = o;
    };
    int p = p0 + lane_p;
    for (; p + 3 * m.lanes < p1; p += 4 * m.lanes) {   // four independent 128-bit loads in flight per thread
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(x + base + (long long)p * C));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(x + base + (long long)(p + m.lanes) * C));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(x + base + (long long)(p + 2 * m.lanes) * C));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(x + base + (long long)(p + 3 * m.lanes) * C));
      emit(u0, p); emit(u1, p + m.lanes); emit(u2, p + 2 * m.lanes); emit(u3, p + 3 * m.lanes);
    }
    for (; p < p1; p += m.lanes) emit(__ldg(reinterpret_cast<const uint4*>(x + base + (long long)p * C)), p);
  }
}

cudaError_t launch_groupnorm(const bf16* x, bf16* y, const float* gamma, const float* beta, int B, int HW, int C, int G,
                             float eps, bool silu, float* workspace, cudaStream_t stream) {
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > kGnThreads * kGnMaxSlots) return cudaErrorInvalidValue;
  // enough chunks to fill the chip (>= ~4 blocks per SM across the batch), at least 32 pixels per chunk
  int nchunks = (148 * 4 + B - 1) / B;
  if (nchunks > (HW + 31) / 32) nchunks = (HW + 31) / 32;
  if (nchunks > kGnMaxChunks) nchunks = kGnMaxChunks;
  if (nchunks < 1) nchunks = 1;
  dim3 grid(nchunks, B);
  float* stats = workspace + (size_t)B * kGnMaxChunks * G * 2;
  cudaError_t e;
  if (deterministic_mode())
    e = launch_pdl(groupnorm_stats_det_kernel, dim3(nchunks, G, B), dim3(kGnThreads), 0, stream, x, workspace, HW, C, G,
                   nchunks);
  else
    e = launch_pdl(groupnorm_stats_kernel, grid, dim3(kGnThreads), 0, stream, x, workspace, HW, C, G, nchunks);
  if (e != cudaSuccess) return e;
  e = launch_pdl(groupnorm_finalize_kernel, dim3(B), dim3(1024), 0, stream, (const float*)workspace, stats, HW, C, G,
                 nchunks, eps);
  if (e != cudaSuccess) return e;
  return launch_pdl(groupnorm_apply_kernel, grid, dim3(kGnThreads), 0, stream, x, y, gamma, beta, (const float*)stats, HW,
                    C, G, nchunks, silu ? 1 : 0);
}

// Statistics accumulated by the producing GEMM's epilogue (GemmParams::gn_sums): (sum, sum sq) per (image, group)
// -> (mean, rstd); replaces groupnorm_stats_kernel + groupnorm_finalize_kernel (one full read of x less).
__global__ void groupnorm_finalize_sums_kernel(const float* __restrict__ sums, float* __restrict__ stats, int n_bg,
                                               double inv_n, float eps) {
  pdl_wait();
  pdl_trigger();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bg) return;
  const double mean = (double)sums[2 * i] * inv_n;
  double var = (double)sums[2 * i + 1] * inv_n - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

cudaError_t launch_groupnorm_from_sums(const bf16* x, bf16* y, const float* gamma, const float* beta, int B, int HW,
                                       int C, int G, float eps, bool silu, const float* sums, float* workspace,
                                       cudaStream_t stream) {
  if (C % 8 != 0 || C % G != 0 || G > 64 || C / 8 > kGnThreads * kGnMaxSlots) return cudaErrorInvalidValue;
  int nchunks = (148 * 4 + B - 1) / B;
  if (nchunks > (HW + 31) / 32) nchunks = (HW + 31) / 32;
  if (nchunks > kGnMaxChunks) nchunks = kGnMaxChunks;
  if (nchunks < 1) nchunks = 1;
  float* stats = workspace + (size_t)B * kGnMaxChunks * G * 2;
  const int n_bg = B * G;
  cudaError_t e = launch_pdl(groupnorm_finalize_sums_kernel, dim3((n_bg + 127) / 128), dim3(128), 0, stream, sums, stats,
                             n_bg, 1.0 / ((double)HW * (C / G)), eps);
  if (e != cudaSuccess) return e;
  return launch_pdl(groupnorm_apply_kernel, dim3(nchunks, B), dim3(kGnThreads), 0, stream, x, y, gamma, beta,
                    (const float*)stats, HW, C, G, nchunks, silu ? 1 : 0);
}

// ------------------------------------------------------------------------------------------ LayerNorm
// One warp per token row; the row lives in registers between the statistics and the normalisation (one global
// read, one write). kV = 128-bit vectors per lane (C <= 256 * kV).
//
// Round 2: persistent form. The first version launched one warp per row in a single wave (8192 rows = 1024 blocks, all
// resident): every warp of the chip loaded, then reduced, then fetched gamma / beta, then stored AT THE SAME TIME, so
// the phases (load latency + load bandwidth + arithmetic + 10 KB of gamma / beta through L1 per row + store bandwidth)
// added up instead of overlapping: 14.8 us isolated / 20.7 us in the step for the 21 MB L2-resident 8192 x 1280 tensor
// (2.8 TB/s), 180 launches per SDXL step. Now: one block of 16 warps per SM, every warp walks rows
// (warp, warp + total, ...) with the NEXT row's loads in flight while the current one is reduced and stored, and
// gamma / beta come from shared memory (filled before the PDL wait: they are weights, never written by the preceding
// kernel). GDF_LN_V1=1 restores the single-wave kernel (A/B timing).
template <int kV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                 const float* __restrict__ beta, long long M, int C, float eps, const float* __restrict__ mod_scale,
                 const float* __restrict__ mod_shift, int rows_per_batch) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
  const int nv = C / 8;
  uint4 u[kV];
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    u[i] = (v < nv) ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
  }
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    float2 f;
    f = unpack_bf16x2(u[i].x); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u[i].y); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u[i].z); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u[i].w); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float mean = s / C;
  const float var = fmaxf(q / C - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float* ms = nullptr;
  const float* mh = nullptr;
  if (mod_scale) {
    const long long b = row / rows_per_batch;
    ms = mod_scale + b * C;
    mh = mod_shift + b * C;
  }
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int v = lane + i * 32;
    if (v >= nv) continue;
    float t[8];
    float2 f;
    f = unpack_bf16x2(u[i].x); t[0] = f.x; t[1] = f.y;
    f = unpack_bf16x2(u[i].y); t[2] = f.x; t[3] = f.y;
    f = unpack_bf16x2(u[i].z); t[4] = f.x; t[5] = f.y;
    f = unpack_bf16x2(u[i].w); t[6] = f.x; t[7] = f.y;
    float gm[8], bt[8];
    if (gamma) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8) + 1);
      gm[0] = g0.x; gm[1] = g0.y; gm[2] = g0.z; gm[3] = g0.w; gm[4] = g1.x; gm[5] = g1.y; gm[6] = g1.z; gm[7] = g1.w;
      bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = v * 8 + j;
      float r = (t[j] - mean) * rstd;
      if (gamma) r = r * gm[j] + bt[j];
      if (ms) r = r * (1.f + __ldg(ms + c)) + __ldg(mh + c);
      t[j] = r;
    }
    uint4 o;
    o.x = pack_bf16x2(t[0], t[1]);
    o.y = pack_bf16x2(t[2], t[3]);
    o.z = pack_bf16x2(t[4], t[5]);
    o.w = pack_bf16x2(t[6], t[7]);
    yr[v] = o;
  }
}

constexpr int kLnThreads = 512;   // persistent kernel: 16 warps, one block per SM
template <int kV>
__global__ void __launch_bounds__(kLnThreads, 1)
layernorm_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, long long M, int C, float eps,
                      const float* __restrict__ mod_scale, const float* __restrict__ mod_shift, int rows_per_batch) {
  extern __shared__ float4 ln_affine[];   // [C / 4] gamma | [C / 4] beta (only when gamma != null)
  const int lane = threadIdx.x & 31;
  const int nv = C / 8;
  if (gamma) {   // weights: safe to read while the preceding kernel is still running
    for (int i = threadIdx.x; i < C / 4; i += blockDim.x) {
      ln_affine[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      ln_affine[C / 4 + i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
    }
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  const long long wstep = (long long)gridDim.x * (blockDim.x >> 5);
  // rows interleaved over (warp slot, block): consecutive rows go to different SMs, every SM gets M / gridDim rows +- 1
  long long row = (long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x;
  uint4 u[kV], un[kV];
  auto load_row = [&](uint4* dst, long long r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + r * C);
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int v = lane + i * 32;
      dst[i] = (v < nv) ? __ldg(xr + v) : make_uint4(0, 0, 0, 0);
    }
  };
  if (row < M) load_row(u, row);
  for (; row < M; row += wstep) {
    const long long nrow = row + wstep;
    if (nrow < M) load_row(un, nrow);   // in flight while this row is reduced, normalised and stored
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      float2 f;
      f = unpack_bf16x2(u[i].x); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
      f = unpack_bf16x2(u[i].y); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
      f = unpack_bf16x2(u[i].z); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
      f = unpack_bf16x2(u[i].w); s += f.x + f.y; q += f.x * f.x + f.y * f.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float mean = s / C;
    const float var = fmaxf(q / C - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float* ms = nullptr;
    const float* mh = nullptr;
    if (mod_scale) {
      const long long b = row / rows_per_batch;
      ms = mod_scale + b * C;
      mh = mod_shift + b * C;
    }
    uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int v = lane + i * 32;
      if (v < nv) {
        float t[8];
        float2 f;
        f = unpack_bf16x2(u[i].x); t[0] = f.x; t[1] = f.y;
        f = unpack_bf16x2(u[i].y); t[2] = f.x; t[3] = f.y;
        f = unpack_bf16x2(u[i].z); t[4] = f.x; t[5] = f.y;
        f = unpack_bf16x2(u[i].w); t[6] = f.x; t[7] = f.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (t[j] - mean) * rstd;
        if (gamma) {
          const float4 g0 = ln_affine[v * 2], g1 = ln_affine[v * 2 + 1];
          const float4 b0 = ln_affine[C / 4 + v * 2], b1 = ln_affine[C / 4 + v * 2 + 1];
          t[0] = t[0] * g0.x + b0.x; t[1] = t[1] * g0.y + b0.y; t[2] = t[2] * g0.z + b0.z; t[3] = t[3] * g0.w + b0.w;
          t[4] = t[4] * g1.x + b1.x; t[5] = t[5] * g1.y + b1.y; t[6] = t[6] * g1.z + b1.z; t[7] = t[7] * g1.w + b1.w;
        }
        if (ms) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(ms + v * 8));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(ms + v * 8) + 1);
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(mh + v * 8));
          const float4 h1 = __ldg(reinterpret_cast<const float4*>(mh + v * 8) + 1);
          t[0] = t[0] * (1.f + s0.x) + h0.x; t[1] = t[1] * (1.f + s0.y) + h0.y;
          t[2] = t[2] * (1.f + s0.z) + h0.z; t[3] = t[3] * (1.f + s0.w) + h0.w;
          t[4] = t[4] * (1.f + s1.x) + h1.x; t[5] = t[5] * (1.f + s1.y) + h1.y;
          t[6] = t[6] * (1.f + s1.z) + h1.z; t[7] = t[7] * (1.f + s1.w) + h1.w;
        }
        uint4 o;
        o.x = pack_bf16x2(t[0], t[1]);
        o.y = pack_bf16x2(t[2], t[3]);
        o.z = pack_bf16x2(t[4], t[5]);
        o.w = pack_bf16x2(t[6], t[7]);
        yr[v] = o;
      }
    }
#pragma unroll
    for (int i = 0; i < kV; ++i) u[i] = un[i];
  }
}

template <int kV>
static cudaError_t launch_layernorm_rows(const bf16* x, bf16* y, const float* gamma, const float* beta, long long M, int C,
                                         float eps, const float* mod_scale, const float* mod_shift, int rpb,
                                         cudaStream_t stream) {
  const size_t smem = gamma ? (size_t)C * 8 : 0;   // <= 10 KB (C <= 1280)
  const long long warps = kLnThreads / 32;
  long long blocks = (M + warps - 1) / warps;
  const int sms = gemm_num_sms();
  if (blocks > sms) blocks = sms;
  return launch_pdl(layernorm_rows_kernel<kV>, dim3((unsigned)blocks), dim3(kLnThreads), smem, stream, x, y, gamma, beta,
                    M, C, eps, mod_scale, mod_shift, rpb);
}

cudaError_t launch_layernorm(const bf16* x, bf16* y, const float* gamma, const float* beta, long long M, int C,
                             float eps, const float* mod_scale, const float* mod_shift, int rows_per_batch,
                             cudaStream_t stream) {
  if (C % 8 != 0) return cudaErrorInvalidValue;
  const int rpb = rows_per_batch > 0 ? rows_per_batch : 1;
  const int nv = C / 8;
  static const bool v1 = [] { const char* e = getenv("GDF_LN_V1"); return e && e[0] == '1'; }();
  // persistent kernel: needs 16-byte aligned affine / modulation vectors (float4 loads) and gamma + beta in 64 KB
  const bool aligned = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                         reinterpret_cast<uintptr_t>(mod_scale) | reinterpret_cast<uintptr_t>(mod_shift)) & 15) == 0;
  // (rows wider than 1280 keep the single-wave kernel: two rows of 12 vectors per lane do not fit 128 registers)
  if (!v1 && aligned && nv <= 160 && (gamma == nullptr) == (beta == nullptr)) {
    if (nv <= 64) return launch_layernorm_rows<2>(x, y, gamma, beta, M, C, eps, mod_scale, mod_shift, rpb, stream);
    if (nv <= 96) return launch_layernorm_rows<3>(x, y, gamma, beta, M, C, eps, mod_scale, mod_shift, rpb, stream);
    return launch_layernorm_rows<5>(x, y, gamma, beta, M, C, eps, mod_scale, mod_shift, rpb, stream);
  }
  const int rows_per_block = 8;
  const long long blocks = (M + rows_per_block - 1) / rows_per_block;
  if (nv <= 64)
    return launch_pdl(layernorm_kernel<2>, dim3((unsigned)blocks), dim3(256), 0, stream, x, y, gamma, beta, M, C, eps,
                      mod_scale, mod_shift, rpb);
  else if (nv <= 96)
    return launch_pdl(layernorm_kernel<3>, dim3((unsigned)blocks), dim3(256), 0, stream, x, y, gamma, beta, M, C, eps,
                      mod_scale, mod_shift, rpb);
  else if (nv <= 160)
    return launch_pdl(layernorm_kernel<5>, dim3((unsigned)blocks), dim3(256), 0, stream, x, y, gamma, beta, M, C, eps,
                      mod_scale, mod_shift, rpb);
  else if (nv <= 384)
    return launch_pdl(layernorm_kernel<12>, dim3((unsigned)blocks), dim3(256), 0, stream, x, y, gamma, beta, M, C, eps,
                      mod_scale, mod_shift, rpb);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace gdf
