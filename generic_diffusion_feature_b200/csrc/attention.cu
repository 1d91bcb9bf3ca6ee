// Flash-style attention forward (online softmax, no N x N materialisation), head_dim 64, bf16 in/out,
// fp32 softmax + accumulation. Reference call site: F.scaled_dot_product_attention in
// feature/diffusers/models/attention_processor.py:3311-3313 (self- and cross-attention, Nk = 77 for text).
// Q/K/V are read in place from the token-major projection outputs (head h = columns [64h, 64h+64)), so the
// q/k/v captures written by the projection GEMM epilogue and the attention input are the same tensors.
//
// v1 data path: cp.async double-buffered K/V tiles in XOR-swizzled shared memory, ldmatrix fragments,
// mma.sync.m16n8k16 (legacy tensor path). The tcgen05/TMEM variant replaces the two mma loops; the softmax
// and pipeline structure stay.
#include <stdlib.h>
#include "ops.h"

namespace gdf {

constexpr int kAttBM = 128;   // query rows per CTA (8 warps x 16)
constexpr int kAttBN = 64;    // kv rows per tile
constexpr int kAttD = 64;
constexpr int kAttThreads = 256;
constexpr int kAttSmem = (kAttBM * kAttD + 4 * kAttBN * kAttD) * 2;  // Q + 2x(K,V) = 48 KB

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile row r, 16-byte chunk c (0..7) -> byte offset inside a [rows][64] bf16 tile with XOR swizzle
__device__ __forceinline__ int swz(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(kAttThreads, 2)
attention64_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, int ldk,
                   const bf16* __restrict__ V, int ldv, bf16* __restrict__ O, int ldo, int Nq, int Nk,
                   float scale_log2) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  uint8_t* sQ = att_smem;
  uint8_t* sK = att_smem + kAttBM * kAttD * 2;
  uint8_t* sV = sK + 2 * kAttBN * kAttD * 2;
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kAttBM;
  const int h = blockIdx.y, b = blockIdx.z;
  const bf16* Qb = Q + ((long long)b * Nq) * ldq + h * kAttD;
  const bf16* Kb = K + ((long long)b * Nk) * ldk + h * kAttD;
  const bf16* Vb = V + ((long long)b * Nk) * ldv + h * kAttD;

  // ---- Q tile (rows beyond Nq are zero-filled)
  for (int i = tid; i < kAttBM * 8; i += kAttThreads) {
    const int r = i >> 3, c = i & 7;
    const int qr = q0 + r;
    const bool ok = qr < Nq;
    cp_async16(sQ + swz(r, c), Qb + (long long)(ok ? qr : 0) * ldq + c * 8, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * kAttBN;
    for (int i = tid; i < kAttBN * 8; i += kAttThreads) {
      const int r = i >> 3, c = i & 7;
      const int kr = k0 + r;
      const bool ok = kr < Nk;
      const long long ro = (long long)(ok ? kr : 0);
      cp_async16(sK + buf * kAttBN * kAttD * 2 + swz(r, c), Kb + ro * ldk + c * 8, ok ? 16 : 0);
      cp_async16(sV + buf * kAttBN * kAttD * 2 + swz(r, c), Vb + ro * ldv + c * 8, ok ? 16 : 0);
    }
  };
  const int ntiles = (Nk + kAttBN - 1) / kAttBN;
  load_kv(0, 0);
  cp_async_commit();

  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ldmatrix_x4(qf[ks], sQ + swz(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)));
    }
    const uint8_t* sKb = sK + buf * kAttBN * kAttD * 2;
    const uint8_t* sVb = sV + buf * kAttBN * kAttD * 2;

    // ---- S = Q K^T  (16 x 64 per warp)
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {   // pairs of 8-wide kv blocks
        uint32_t kf[4];
        ldmatrix_x4(kf, sKb + swz(nb2 * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
        mma_bf16_16816(s_acc[nb2 * 2], qf[ks], kf[0], kf[1]);
        mma_bf16_16816(s_acc[nb2 * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    // ---- scale, mask, online softmax
    const int kv0 = t * kAttBN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = kv0 + nb * 8 + (lane & 3) * 2 + (j & 1);
        float v = s_acc[nb][j] * scale_log2;
        if (col >= Nk) v = -INFINITY;
        s_acc[nb][j] = v;
        mx[j >> 1] = fmaxf(mx[j >> 1], v);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = exp2f(s_acc[nb][0] - m_run[0]);
      const float p1 = exp2f(s_acc[nb][1] - m_run[0]);
      const float p2 = exp2f(s_acc[nb][2] - m_run[1]);
      const float p3 = exp2f(s_acc[nb][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      // A fragment of P for k-step nb/2: a0,a1 from even block, a2,a3 from odd block
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      o_acc[nb][0] *= alpha[0];
      o_acc[nb][1] *= alpha[0];
      o_acc[nb][2] *= alpha[1];
      o_acc[nb][3] *= alpha[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {       // 16 kv rows per step
#pragma unroll
      for (int db2 = 0; db2 < 4; ++db2) {  // pairs of 8-wide d blocks
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, sVb + swz(ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), db2 * 2 + (lane >> 4)));
        mma_bf16_16816(o_acc[db2 * 2], pf[ks], vf[0], vf[1]);
        mma_bf16_16816(o_acc[db2 * 2 + 1], pf[ks], vf[2], vf[3]);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled
  }

  // ---- finalise: O / l, stage through this warp's rows of sQ, coalesced 16 B stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = warp * 16 + (lane >> 2);
#pragma unroll
  for (int nb = 0; nb < 8; ++nb) {
    const int colb = ((lane & 3) * 2) * 2;  // byte offset inside the 16 B chunk
    *reinterpret_cast<uint32_t*>(sQ + swz(r0, nb) + colb) = pack_bf16x2(o_acc[nb][0] * inv0, o_acc[nb][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + swz(r0 + 8, nb) + colb) = pack_bf16x2(o_acc[nb][2] * inv1, o_acc[nb][3] * inv1);
  }
  __syncwarp();
  bf16* Ob = O + ((long long)b * Nq) * ldo + h * kAttD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int r = warp * 16 + (idx >> 3), c = idx & 7;
    const int qr = q0 + r;
    if (qr < Nq) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + swz(r, c));
      *reinterpret_cast<uint4*>(Ob + (long long)qr * ldo + c * 8) = v;
    }
  }
}

static bool attention_tc_enabled() {   // GDF_ATTN_TC=0: round-1 kernels (A/B timing)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GDF_ATTN_TC");
    v = e ? atoi(e) : 1;
  }
  return v != 0;
}

bool attention_uses_tcgen05(int Nk) {
  // long KV (self-attention): the QKV projection writes V as fp16 and P.V runs in fp16; short KV (text cross-attention,
  // Nk = 77): V stays bf16 (one batched projection for all blocks) and P is rounded to bf16
  static int use_tc = -1;
  if (use_tc < 0) {
    const char* v = getenv("GDF_ATTN_TCGEN05");
    use_tc = v ? atoi(v) : 1;
  }
  return use_tc && Nk >= 128;
}

cudaError_t launch_attention64(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo,
                               int B, int heads, int Nq, int Nk, float scale, int v_f16, cudaStream_t stream) {
  if ((ldq | ldk | ldv | ldo) % 8 != 0 || Nk < 1) return cudaErrorInvalidValue;
  if (attention_tc_enabled()) {
    if (launch_attention_tc(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, 64, scale, v_f16, nullptr, stream) != 0) {
      fprintf(stderr, "gdf: tcgen05 attention launch failed: %s\n", last_error().c_str());
      return cudaErrorUnknown;
    }
    return cudaSuccess;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (v_f16) {
    if (!attention_uses_tcgen05(Nk)) return cudaErrorInvalidValue;   // the mma.sync kernel takes bf16 V
    if (launch_attention64_tcgen05(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, scale, stream) != 0) {
      fprintf(stderr, "gdf: tcgen05 attention launch failed: %s\n", last_error().c_str());
      return cudaErrorUnknown;
    }
    return cudaSuccess;
  }
  dim3 grid((Nq + kAttBM - 1) / kAttBM, heads, B);
  return launch_pdl(attention64_kernel, grid, dim3(kAttThreads), kAttSmem, stream, Q, ldq, K, ldk, V, ldv, O, ldo, Nq, Nk,
                    scale * 1.4426950408889634f);
}

// ------------------------------------------------------------------------------------------ row softmax
// In-place softmax over bf16 rows (VAE mid-block single-head attention, d = 512: scores are materialised by the
// batched GEMM, normalised here, then multiplied with V by a second GEMM).
__global__ void __launch_bounds__(256)
softmax_rows_kernel(bf16* __restrict__ S, long long rows, int cols, int ld) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  if (row >= rows) return;
  bf16* s = S + row * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = cols / 8;
  float mx = -INFINITY;
  for (int v = tid; v < nv; v += 256) {
    const uint4 u = reinterpret_cast<const uint4*>(s)[v];
    float2 f;
    f = unpack_bf16x2(u.x); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.y); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.z); mx = fmaxf(mx, fmaxf(f.x, f.y));
    f = unpack_bf16x2(u.w); mx = fmaxf(mx, fmaxf(f.x, f.y));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int v = tid; v < nv; v += 256) {
    const uint4 u = reinterpret_cast<const uint4*>(s)[v];
    float2 f;
    f = unpack_bf16x2(u.x); sum += __expf(f.x - mx) + __expf(f.y - mx);
    f = unpack_bf16x2(u.y); sum += __expf(f.x - mx) + __expf(f.y - mx);
    f = unpack_bf16x2(u.z); sum += __expf(f.x - mx) + __expf(f.y - mx);
    f = unpack_bf16x2(u.w); sum += __expf(f.x - mx) + __expf(f.y - mx);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int v = tid; v < nv; v += 256) {
    uint4 u = reinterpret_cast<const uint4*>(s)[v];
    float2 f;
    f = unpack_bf16x2(u.x); u.x = pack_bf16x2(__expf(f.x - mx) * inv, __expf(f.y - mx) * inv);
    f = unpack_bf16x2(u.y); u.y = pack_bf16x2(__expf(f.x - mx) * inv, __expf(f.y - mx) * inv);
    f = unpack_bf16x2(u.z); u.z = pack_bf16x2(__expf(f.x - mx) * inv, __expf(f.y - mx) * inv);
    f = unpack_bf16x2(u.w); u.w = pack_bf16x2(__expf(f.x - mx) * inv, __expf(f.y - mx) * inv);
    reinterpret_cast<uint4*>(s)[v] = u;
  }
}

cudaError_t launch_softmax_rows(bf16* S, long long rows, int cols, int ld, cudaStream_t stream) {
  if (cols % 8 != 0 || ld % 8 != 0 || rows > 0x7fffffffLL) return cudaErrorInvalidValue;
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(S, rows, cols, ld);
  return cudaGetLastError();
}

// ------------------------------------------------------- tensor-core form of the probability-map path (self maps)
// P = softmax(S) for fp32 scores S (written by the batched tcgen05 GEMM of Q K^T with alpha = scale) -> fp16
// probabilities, the `...-self-map` feature itself. One block per row; the row is read twice from L2 (max + sum of
// exponentials in one pass over registers would need cols / 256 live values: <= 16 for 4096 keys, kept in registers).
constexpr int kSmThreads = 256, kSmMaxPer = 32;   // rows of up to 8192 keys
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_f32_f16_kernel(const float* __restrict__ S, __half* __restrict__ P, long long rows, int cols) {
  __shared__ float red[kSmThreads / 32];
  const long long row = blockIdx.x;
  if (row >= rows) return;
  const float* s = S + row * cols;
  __half* p = P + row * cols;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float v[kSmMaxPer];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    const int c = tid + i * kSmThreads;
    v[i] = c < cols ? __ldcs(s + c) : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kSmThreads / 32; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    v[i] = __expf(v[i] - mx);     // exp(-inf) = 0 for the padding lanes
    sum += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSmThreads / 32; ++i) sum += red[i];
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < kSmMaxPer; ++i) {
    const int c = tid + i * kSmThreads;
    if (c < cols) p[c] = __float2half_rn(v[i] * inv);
  }
}
cudaError_t launch_softmax_rows_f32_f16(const float* S, __half* P, long long rows, int cols, cudaStream_t stream) {
  if (cols < 1 || cols > kSmThreads * kSmMaxPer || rows > 0x7fffffffLL) return cudaErrorInvalidValue;
  softmax_rows_f32_f16_kernel<<<(unsigned)rows, kSmThreads, 0, stream>>>(S, P, rows, cols);
  return cudaGetLastError();
}

// V of one image, token-major bf16 [Nk, heads * D] (row pitch ldv) -> V^T fp16 [heads][D][Nk]: the K-major B operand
// of the P V product on the fp16 GEMM path (P is fp16).
__global__ void __launch_bounds__(256)
transpose_v_f16_kernel(const bf16* __restrict__ V, int ldv, __half* __restrict__ VT, int Nk, int D) {
  __shared__ float tile[32][33];
  const int h = blockIdx.z;
  const int k0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, d = d0 + tx;
    tile[r][tx] = (k < Nk && d < D) ? __bfloat162float(V[(long long)k * ldv + h * D + d]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int d = d0 + r, k = k0 + tx;
    if (d < D && k < Nk) VT[((long long)h * D + d) * Nk + k] = __float2half_rn(tile[tx][r]);
  }
}
cudaError_t launch_transpose_v_f16(const bf16* V, int ldv, __half* VT, int heads, int Nk, int D, cudaStream_t stream) {
  const dim3 grid((Nk + 31) / 32, (D + 31) / 32, heads);
  transpose_v_f16_kernel<<<grid, 256, 0, stream>>>(V, ldv, VT, Nk, D);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- attention with the probability matrix as an output
// Slow path of the reference (AttnStoreProcessor, feature/components/attention.py:165-263): P = softmax(scale Q K^T)
// is materialised per head - it is the `...-self-map` / `...-cross-map` feature, (B, heads, Nq, Nk), and its head mean
// feeds the AttentionStore - then O = P V. CUDA cores, fp32, two passes over the keys (row maximum / sum, then
// probabilities + P V); one CTA = 16 query rows of one (batch, head), one thread = one key of a 128-key tile.
constexpr int kApQT = 16, kApKT = 128;

__device__ __forceinline__ void ap_scores(const float* __restrict__ q_s, const bf16* __restrict__ krow, int D, float* acc) {
#pragma unroll
  for (int q = 0; q < kApQT; ++q) acc[q] = 0.f;
  for (int d = 0; d < D; d += 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(krow + d);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    float kf[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(b2[i]);
      kf[2 * i] = f.x;
      kf[2 * i + 1] = f.y;
    }
#pragma unroll
    for (int q = 0; q < kApQT; ++q) {
      const float* qr = q_s + q * D + d;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[q] = fmaf(qr[i], kf[i], acc[q]);
    }
  }
}

__global__ void __launch_bounds__(kApKT)
attention_probs_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, int ldk,
                       const bf16* __restrict__ V, int ldv, int v_f16, bf16* __restrict__ O, int ldo,
                       __half* __restrict__ P, int heads, int Nq, int Nk, int D, float scale,
                       const float* __restrict__ key_bias, __half* __restrict__ P2, int split) {
  extern __shared__ float ap_smem[];
  float* q_s = ap_smem;                       // [16][D] (pre-scaled queries)
  float* p_s = ap_smem + kApQT * D;           // [16][128] probabilities of the current key tile / reduction scratch
  __shared__ float m_s[kApQT], l_s[kApQT];
  const int tid = threadIdx.x;
  const int q0 = blockIdx.x * kApQT, h = blockIdx.y, b = blockIdx.z;
  for (int i = tid; i < kApQT * D; i += kApKT) {
    const int q = i / D, d = i - q * D, row = q0 + q;
    q_s[i] = row < Nq ? __bfloat162float(Q[((long long)b * Nq + row) * ldq + h * D + d]) * scale : 0.f;
  }
  __syncthreads();
  const bf16* Kb = K + (long long)b * Nk * ldk + h * D;
  const bf16* Vb = V + (long long)b * Nk * ldv + h * D;
  // ---- pass 1: per-thread running (max, sum) over this thread's keys, then a block reduction per query row
  float m_t[kApQT], l_t[kApQT], acc[kApQT];
#pragma unroll
  for (int q = 0; q < kApQT; ++q) { m_t[q] = -INFINITY; l_t[q] = 0.f; }
  for (int k0 = 0; k0 < Nk; k0 += kApKT) {
    const int j = k0 + tid;
    if (j < Nk) {
      ap_scores(q_s, Kb + (long long)j * ldk, D, acc);
      if (key_bias) {   // additive key mask in logit units (PixArt: (1 - mask) * -10000, attention.py:165-263 path)
        const float kb = __ldg(key_bias + (long long)b * Nk + j);
#pragma unroll
        for (int q = 0; q < kApQT; ++q) acc[q] += kb;
      }
#pragma unroll
      for (int q = 0; q < kApQT; ++q) {
        const float mn = fmaxf(m_t[q], acc[q]);
        l_t[q] = l_t[q] * __expf(m_t[q] - mn) + __expf(acc[q] - mn);
        m_t[q] = mn;
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kApQT; ++q) p_s[q * kApKT + tid] = m_t[q];
  __syncthreads();
  if (tid < kApQT) {
    float m = -INFINITY;
    for (int i = 0; i < kApKT; ++i) m = fmaxf(m, p_s[tid * kApKT + i]);
    m_s[tid] = m;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kApQT; ++q) p_s[q * kApKT + tid] = (l_t[q] > 0.f) ? l_t[q] * __expf(m_t[q] - m_s[q]) : 0.f;
  __syncthreads();
  if (tid < kApQT) {
    float l = 0.f;
    for (int i = 0; i < kApKT; ++i) l += p_s[tid * kApKT + i];
    l_s[tid] = l;
  }
  __syncthreads();
  // ---- pass 2: probabilities (written out) and O = P V; thread tid owns output dims tid and tid + 128
  float o0[kApQT], o1[kApQT];
#pragma unroll
  for (int q = 0; q < kApQT; ++q) { o0[q] = 0.f; o1[q] = 0.f; }
  const int d0 = tid, d1 = tid + kApKT;
  for (int k0 = 0; k0 < Nk; k0 += kApKT) {
    const int j = k0 + tid;
    __syncthreads();                        // p_s of the previous tile has been consumed
    if (j < Nk) {
      ap_scores(q_s, Kb + (long long)j * ldk, D, acc);
      if (key_bias) {
        const float kb = __ldg(key_bias + (long long)b * Nk + j);
#pragma unroll
        for (int q = 0; q < kApQT; ++q) acc[q] += kb;
      }
#pragma unroll
      for (int q = 0; q < kApQT; ++q) {
        const float pr = __expf(acc[q] - m_s[q]) / l_s[q];
        p_s[q * kApKT + tid] = pr;
        if (q0 + q < Nq) {
          // keys >= split go to P (pitch Nk - split), keys < split to P2 (pitch split): the Flux joint attention stores
          // image-query x text-key (`cross-map`) and image x image (`self-map`) separately; split = 0 elsewhere
          const long long rowi = ((long long)b * heads + h) * Nq + q0 + q;
          if (j >= split) {
            if (P) P[rowi * (Nk - split) + (j - split)] = __float2half_rn(pr);
          } else if (P2) {
            P2[rowi * split + j] = __float2half_rn(pr);
          }
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < kApQT; ++q) p_s[q * kApKT + tid] = 0.f;
    }
    __syncthreads();
    const int nk = min(kApKT, Nk - k0);
    for (int jj = 0; jj < nk; ++jj) {
      const long long vo = (long long)(k0 + jj) * ldv;
      float v0 = 0.f, v1 = 0.f;
      if (v_f16) {
        const __half* vh = reinterpret_cast<const __half*>(Vb);
        if (d0 < D) v0 = __half2float(vh[vo + d0]);
        if (d1 < D) v1 = __half2float(vh[vo + d1]);
      } else {
        if (d0 < D) v0 = __bfloat162float(Vb[vo + d0]);
        if (d1 < D) v1 = __bfloat162float(Vb[vo + d1]);
      }
#pragma unroll
      for (int q = 0; q < kApQT; ++q) {
        const float pr = p_s[q * kApKT + jj];
        o0[q] = fmaf(pr, v0, o0[q]);
        o1[q] = fmaf(pr, v1, o1[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < kApQT; ++q) {
    const int row = q0 + q;
    if (row < Nq) {
      bf16* orow = O + ((long long)b * Nq + row) * ldo + h * D;
      if (d0 < D) orow[d0] = __float2bfloat16_rn(o0[q]);
      if (d1 < D) orow[d1] = __float2bfloat16_rn(o1[q]);
    }
  }
}
cudaError_t launch_attention_probs(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, int v_f16,
                                   bf16* O, int ldo, __half* P, int B, int heads, int Nq, int Nk, int D, float scale,
                                   cudaStream_t stream, const float* key_bias, __half* P2, int split) {
  if (D % 8 != 0 || D > 2 * kApKT || (ldq | ldk | ldv) % 8 != 0 || Nk < 1 || Nq < 1) return cudaErrorInvalidValue;
  const dim3 grid((Nq + kApQT - 1) / kApQT, heads, B);
  const size_t smem = (size_t)(kApQT * D + kApQT * kApKT) * sizeof(float);
  attention_probs_kernel<<<grid, kApKT, smem, stream>>>(Q, ldq, K, ldk, V, ldv, v_f16, O, ldo, P, heads, Nq, Nk, D, scale,
                                                        key_bias, P2, split);
  return cudaGetLastError();
}

// head mean of a probability map: P fp16 [B, heads, n] -> mean fp16 [B, n] (AttnStoreProcessor: to_store.mean(1))
__global__ void head_mean_kernel(const __half* __restrict__ P, __half* __restrict__ out, int heads, long long n,
                                 long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / n, r = i - b * n;
    float s = 0.f;
    for (int hh = 0; hh < heads; ++hh) s += __half2float(P[(b * heads + hh) * n + r]);
    out[i] = __float2half_rn(s / (float)heads);
  }
}
cudaError_t launch_head_mean(const __half* P, __half* out, int B, int heads, long long n, cudaStream_t stream) {
  const long long total = (long long)B * n;
  const long long blocks = (total + 255) / 256;
  head_mean_kernel<<<(unsigned)(blocks < 148 * 32 ? (blocks < 1 ? 1 : blocks) : 148 * 32), 256, 0, stream>>>(
      P, out, heads, n, total);
  return cudaGetLastError();
}

}  // namespace gdf
