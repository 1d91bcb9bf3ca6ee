// Flash-style attention forward for head dims other than 64 (SD-1.5: 40 / 80 / 160, PixArt: 72): the mma.sync
// kernel of attention.cu templated on the padded head dim DP (multiple of 16; columns [D, DP) are zero-filled in
// shared memory, so they add nothing to QK^T and produce zero output columns that are never stored).
// Reference call site: F.scaled_dot_product_attention, attention_processor.py:3311-3313.
#include <stdlib.h>
#include "ops.h"

namespace gdf {

constexpr int kGaBM = 128, kGaBN = 64, kGaThreads = 256;

__device__ __forceinline__ void ga_cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void ga_ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ga_ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ga_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DP>
struct GaCfg {
  static constexpr int kRS = DP * 2 + 16;        // padded row stride in bytes: (kRS / 16) is odd -> conflict-free
  static constexpr int kChunks = DP / 8;         // 16-byte chunks per (padded) row
  static constexpr int kSmem = (kGaBM + 4 * kGaBN) * kRS;
};

template <int DP>
__global__ void __launch_bounds__(kGaThreads, 1)
attention_generic_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, int ldk,
                         const bf16* __restrict__ V, int ldv, bf16* __restrict__ O, int ldo, int Nq, int Nk, int D,
                         float scale_log2, const float* __restrict__ key_bias) {
  using Cfg = GaCfg<DP>;
  constexpr int RS = Cfg::kRS;
  extern __shared__ __align__(128) uint8_t ga_smem[];
  uint8_t* sQ = ga_smem;
  uint8_t* sK = ga_smem + kGaBM * RS;
  uint8_t* sV = sK + 2 * kGaBN * RS;
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kGaBM;
  const int h = blockIdx.y, b = blockIdx.z;
  const bf16* Qb = Q + ((long long)b * Nq) * ldq + h * D;
  const bf16* Kb = K + ((long long)b * Nk) * ldk + h * D;
  const bf16* Vb = V + ((long long)b * Nk) * ldv + h * D;
  const int real_chunks = D / 8;
  const float* kb = key_bias ? key_bias + (long long)b * Nk : nullptr;

  for (int i = tid; i < kGaBM * Cfg::kChunks; i += kGaThreads) {
    const int r = i / Cfg::kChunks, c = i % Cfg::kChunks;
    const int qr = q0 + r;
    const bool ok = qr < Nq && c < real_chunks;
    ga_cp_async16(sQ + r * RS + c * 16, Qb + (long long)(ok ? qr : 0) * ldq + (ok ? c : 0) * 8, ok ? 16 : 0);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * kGaBN;
    for (int i = tid; i < kGaBN * Cfg::kChunks; i += kGaThreads) {
      const int r = i / Cfg::kChunks, c = i % Cfg::kChunks;
      const int kr = k0 + r;
      const bool ok = kr < Nk && c < real_chunks;
      const long long ro = ok ? kr : 0;
      const int co = ok ? c * 8 : 0;
      ga_cp_async16(sK + (buf * kGaBN + r) * RS + c * 16, Kb + ro * ldk + co, ok ? 16 : 0);
      ga_cp_async16(sV + (buf * kGaBN + r) * RS + c * 16, Vb + ro * ldv + co, ok ? 16 : 0);
    }
  };
  const int ntiles = (Nk + kGaBN - 1) / kGaBN;
  load_kv(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  float o_acc[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[DP / 16][4];

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, buf ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < DP / 16; ++ks)
        ga_ldmatrix_x4(qf[ks], sQ + (warp * 16 + (lane & 15)) * RS + (ks * 2 + (lane >> 4)) * 16);
    }
    const uint8_t* sKb = sK + buf * kGaBN * RS;
    const uint8_t* sVb = sV + buf * kGaBN * RS;
    float s_acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s_acc[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < DP / 16; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < 4; ++nb2) {
        uint32_t kf[4];
        ga_ldmatrix_x4(kf, sKb + (nb2 * 16 + (lane & 7) + ((lane >> 4) << 3)) * RS + (ks * 2 + ((lane >> 3) & 1)) * 16);
        ga_mma(s_acc[nb2 * 2], qf[ks], kf[0], kf[1]);
        ga_mma(s_acc[nb2 * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    const int kv0 = t * kGaBN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = kv0 + nb * 8 + (lane & 3) * 2 + (j & 1);
        float v = s_acc[nb][j] * scale_log2;
        if (col >= Nk) v = -INFINITY;
        else if (kb) v = fmaf(__ldg(kb + col), 1.4426950408889634f, v);
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
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int nb = 0; nb < DP / 8; ++nb) {
      o_acc[nb][0] *= alpha[0];
      o_acc[nb][1] *= alpha[0];
      o_acc[nb][2] *= alpha[1];
      o_acc[nb][3] *= alpha[1];
    }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int db2 = 0; db2 < DP / 16; ++db2) {
        uint32_t vf[4];
        ga_ldmatrix_x4_trans(vf, sVb + (ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3)) * RS + (db2 * 2 + (lane >> 4)) * 16);
        ga_mma(o_acc[db2 * 2], pf[ks], vf[0], vf[1]);
        ga_mma(o_acc[db2 * 2 + 1], pf[ks], vf[2], vf[3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = warp * 16 + (lane >> 2);
#pragma unroll
  for (int nb = 0; nb < DP / 8; ++nb) {
    const int colb = nb * 16 + (lane & 3) * 4;
    *reinterpret_cast<uint32_t*>(sQ + r0 * RS + colb) = pack_bf16x2(o_acc[nb][0] * inv0, o_acc[nb][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + (r0 + 8) * RS + colb) = pack_bf16x2(o_acc[nb][2] * inv1, o_acc[nb][3] * inv1);
  }
  __syncwarp();
  bf16* Ob = O + ((long long)b * Nq) * ldo + h * D;
  for (int idx = lane; idx < 16 * real_chunks; idx += 32) {
    const int r = warp * 16 + idx / real_chunks, c = idx % real_chunks;
    const int qr = q0 + r;
    if (qr < Nq) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + r * RS + c * 16);
      *reinterpret_cast<uint4*>(Ob + (long long)qr * ldo + c * 8) = v;
    }
  }
}

template <int DP>
static cudaError_t launch_ga(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo,
                             int B, int heads, int Nq, int Nk, int D, float scale, cudaStream_t stream,
                             const float* key_bias) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_generic_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         GaCfg<DP>::kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((Nq + kGaBM - 1) / kGaBM, heads, B);
  return launch_pdl(attention_generic_kernel<DP>, grid, dim3(kGaThreads), (size_t)GaCfg<DP>::kSmem, stream, Q, ldq, K, ldk,
                    V, ldv, O, ldo, Nq, Nk, D, scale * 1.4426950408889634f, key_bias);
}

cudaError_t launch_attention_generic(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O,
                                     int ldo, int B, int heads, int Nq, int Nk, int D, float scale,
                                     cudaStream_t stream, const float* key_bias) {
  if (D % 8 != 0 || D > 160 || (ldq | ldk | ldv | ldo) % 8 != 0 || Nk < 1) return cudaErrorInvalidValue;
  {
    // tcgen05 / TMEM kernel for every head dim it is built for (GDF_ATTN_TC=0: the mma.sync kernels below, kept for
    // head dim 160 = the 16x16 / 8x8-token levels of SD-1.5, and for A/B timing)
    static int tc = -1;
    if (tc < 0) {
      const char* e = getenv("GDF_ATTN_TC");
      tc = e ? atoi(e) : 1;
    }
    if (tc && attention_tc_supports(D)) {
      if (launch_attention_tc(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, D, scale, 0, key_bias, stream) != 0) {
        fprintf(stderr, "gdf: tcgen05 attention launch failed: %s\n", last_error().c_str());
        return cudaErrorUnknown;
      }
      return cudaSuccess;
    }
  }
  if (D <= 48) return launch_ga<48>(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, D, scale, stream, key_bias);
  if (D <= 80) return launch_ga<80>(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, D, scale, stream, key_bias);
  if (D <= 128) return launch_ga<128>(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, D, scale, stream, key_bias);   // Flux
  return launch_ga<160>(Q, ldq, K, ldk, V, ldv, O, ldo, B, heads, Nq, Nk, D, scale, stream, key_bias);
}

}  // namespace gdf
