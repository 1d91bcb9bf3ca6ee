// Flash attention forward on tcgen05 / TMEM / TMA, head_dim 64 (self-attention of the UNet levels).
// Reference call site: F.scaled_dot_product_attention, feature/diffusers/models/attention_processor.py:3311-3313.
//
// One CTA = 2 query tiles of 128 rows (256 queries of one (batch, head)), looping over KV tiles of 128 rows:
//   warp 0     : TMA producer  (Q tiles once; K and V tiles through two 3-deep rings; 128B swizzle)
//   warp 1     : MMA issuer    (S_t = Q_t K_j^T : UMMA 128x128x16 x4, K-major operands;
//                               PV_t = P_t V_j  : UMMA 128x64x16 x8, A = P (bf16, written to smem by the softmax
//                               warps in the canonical K-major SW128 layout), B = V tile as MN-major operand)
//   warp 2     : TMEM allocator (512 columns: S_A, S_B 128 each; PV_A, PV_B 64 each)
//   warps 4-11 : softmax       (2 groups x 4 warps; thread = one query row: two passes over S in TMEM
//                               (row max, then exp2 / row sum / bf16 P -> smem), running rescale of the fp32
//                               output accumulator kept in registers, PV partial products read back from TMEM)
// Q/K/V are read in place from the token-major projection output (head h = columns [64h, 64h+64)).
#include <stdlib.h>
#include "ops.h"

namespace gdf {

constexpr int kFaThreads = 384;
constexpr int kFaRing = 3;
constexpr int kFaTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16
// smem: Q 2 tiles | K ring | V ring | P 2 x (2 blocks of 16 KB) | barriers
constexpr int kFaOffK = 2 * kFaTile;
constexpr int kFaOffV = kFaOffK + kFaRing * kFaTile;
constexpr int kFaOffP = kFaOffV + kFaRing * kFaTile;
constexpr int kFaOffBar = kFaOffP + 4 * kFaTile;
constexpr int kFaSmem = kFaOffBar + 256 + 1024;

struct FaParams {
  int Nq, Nk, heads;
  int num_kv_tiles;
  float scale_log2;
  bf16* O;
  int ldo;
};

// V = 1: two passes over S in TMEM, fp32 output accumulator in registers (kept for A/B timing: GDF_FA_V1=1).
// V = 2: S read from TMEM once (128 registers, S_t released to the tensor core right away so that S_t(j+1) runs under
//        softmax(j)), output accumulator left in TMEM (P V accumulates over the KV tiles), running maximum updated
//        lazily: O / l are rescaled only when the block maximum exceeds the one in use by more than 2^8 (P <= 256 in
//        fp16), which after the first tiles is rare -> no per-tile read-modify of the accumulator.
// kPingPong: the two softmax groups (one warp of each per SM sub-partition) take turns in the exponential phase
//        through a pair of named barriers, so that one group's MUFU-bound phase runs against the other's TMEM loads /
//        row maxima / barrier round trips instead of against its exponentials (both groups in the MUFU phase at once
//        halve each other's rate and leave the remaining phases uncovered).
template <int V, int kPolyMod, bool kPingPong>
__global__ void __launch_bounds__(kFaThreads, 1)
attention64_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                           const __grid_constant__ CUtensorMap map_v, const FaParams p) {
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFaOffBar);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [ring]
  uint64_t* k_empty = k_full + kFaRing;
  uint64_t* v_full = k_empty + kFaRing;
  uint64_t* v_empty = v_full + kFaRing;
  uint64_t* s_full = v_empty + kFaRing;    // [2] per query tile
  uint64_t* p_full = s_full + 2;           // [2]
  uint64_t* pv_full = p_full + 2;          // [2]
  uint64_t* s_free = pv_full + 2;          // [2] S_t(j) fully read from TMEM (next S may overwrite it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = p.num_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kFaRing; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 4);
      mbar_init(&pv_full[t], 1);
      mbar_init(&s_free[t], 4);
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
  pdl_wait();      // prologue above overlaps the previous kernel's tail; Q/K/V are only read from here on
  pdl_trigger();

  // register rebalancing: the producer / MMA / allocator warpgroup needs few registers, the softmax warpgroups many
  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");

  if (warp == 0) {
    // ================================================= TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kFaTile);
      tma_load_3d(smem, &map_q, q_full, h * 64, q0, b);
      tma_load_3d(smem + kFaTile, &map_q, q_full, h * 64, q0 + 128, b);
    }
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < n; ++j) {
      mbar_wait(&k_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[s], kFaTile);
        tma_load_3d(smem + kFaOffK + s * kFaTile, &map_k, &k_full[s], h * 64, j * 128, b);
      }
      mbar_wait(&v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[s], kFaTile);
        tma_load_3d(smem + kFaOffV + s * kFaTile, &map_v, &v_full[s], h * 64, j * 128, b);
      }
      __syncwarp();
      if (++s == kFaRing) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ================================================= MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0);
    const uint32_t idesc_pv = umma_idesc_f16(128, 64, 1);   // A = P (fp16), B = V tile (fp16), MN-major   // A = P (fp16), B = V tile (bf16), MN-major
    const uint32_t q_addr = smem_u32(smem);
    const uint32_t k_addr = smem_u32(smem + kFaOffK);
    const uint32_t v_addr = smem_u32(smem + kFaOffV);
    const uint32_t p_addr = smem_u32(smem + kFaOffP);
    auto issue_s = [&](int t, int slot) {   // S_t = Q_t K^T into TMEM columns [t*128, t*128+128)
      const uint64_t da = umma_desc_kmajor_sw128(q_addr + t * kFaTile);
      const uint64_t db = umma_desc_kmajor_sw128(k_addr + slot * kFaTile);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + t * 128, da + 2 * k, db + 2 * k, idesc_s, k != 0);
      umma_commit(&s_full[t]);
    };
    auto issue_pv = [&](int t, int slot, bool first) {  // PV_t (+)= P_t V into TMEM columns [256 + t*64, +64)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // A: P block (k / 4) of 16 KB, 32 B step inside the swizzle row; B: 16 kv rows = 2048 B per step
        const uint64_t da = umma_desc_kmajor_sw128(p_addr + t * 2 * kFaTile + (k >> 2) * kFaTile) + 2 * (k & 3);
        const uint64_t db = umma_desc_mnmajor_sw128(v_addr + slot * kFaTile + k * 2048, 8192);
        umma_f16_ss(tmem_base + 256 + t * 64, da, db, idesc_pv, (k != 0 || (V == 2 && !first)) ? 1u : 0u);
      }
      umma_commit(&pv_full[t]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {   // elect.sync: single active lane known to the compiler -> plain uniform-register operands
      issue_s(0, 0);
      issue_s(1, 0);
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    // Event-driven issue: each query tile advances on its own barriers (S_t(j+1) once S_t(j) has been read out of
    // TMEM, PV_t(j) once P_t(j) is in smem), so one group never waits for the other group's softmax.
    int next_s[2] = {1, 1}, next_pv[2] = {0, 0};
    int s_issued[kFaRing] = {0, 0, 0}, pv_issued[kFaRing] = {0, 0, 0};
    while (next_pv[0] < n || next_pv[1] < n) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (next_s[t] < n) {
          const int js = next_s[t], slot = js % kFaRing;
          // warp-uniform decision (a completed phase observed by any lane is complete for all)
          if (__any_sync(0xffffffffu, mbar_try_wait(&s_free[t], (js - 1) & 1) &&
                                          mbar_try_wait(&k_full[slot], (js / kFaRing) & 1))) {
            tc_fence_after();
            if (elect_one()) {
              issue_s(t, slot);
              if (s_issued[slot] == 1) umma_commit(&k_empty[slot]);   // both query tiles have consumed K(js)
            }
            s_issued[slot] ^= 1;
            __syncwarp();
            next_s[t] = js + 1;
          }
        }
        if (next_pv[t] < n) {
          const int jp = next_pv[t], slot = jp % kFaRing;
          if (__any_sync(0xffffffffu, mbar_try_wait(&p_full[t], jp & 1) &&
                                          mbar_try_wait(&v_full[slot], (jp / kFaRing) & 1))) {
            tc_fence_after();
            if (elect_one()) {
              issue_pv(t, slot, jp == 0);
              if (pv_issued[slot] == 1) umma_commit(&v_empty[slot]);
            }
            pv_issued[slot] ^= 1;
            __syncwarp();
            next_pv[t] = jp + 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================= softmax + output accumulation
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int e = warp - 4;
    const int t = e >> 2;        // query tile of this group
    const int quad = e & 3;      // == warp % 4: TMEM lane quadrant
    const int r = quad * 32 + lane;              // row inside the query tile
    const int qrow = q0 + t * 128 + r;
    const uint32_t t_s = tmem_base + (uint32_t(quad * 32) << 16) + t * 128;
    const uint32_t t_pv = tmem_base + (uint32_t(quad * 32) << 16) + 256 + t * 64;
    uint8_t* p_base = smem + kFaOffP + t * 2 * kFaTile;
    if constexpr (V == 2) {
    float m_run = -INFINITY, l_run = 0.f;
    if (kPingPong && t == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");   // group 0 goes first
    const uint32_t p_row = smem_u32(p_base) + r * 128;   // this thread's 128 B row of the P blocks (SW128 K-major)
    const uint32_t p_swz = (r & 7) << 4;
    const float thresh = 8.f / p.scale_log2;   // lazy rescale: keep the maximum in use while P = 2^(..) stays <= 2^8
    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      float sf[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sf[c * 32]));
      tmem_ld_wait();
      // S_t(j) is in registers: the tensor core may overwrite it with S_t(j+1) while this tile's softmax runs
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      const int kv0 = j * 128;
      if (kv0 + 128 > p.Nk) {   // ragged last tile: keys beyond Nk (zero-filled by TMA) are masked out
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kv0 + i >= p.Nk) sf[i] = -INFINITY;
      }
      // ---- row maximum (4 independent chains, 3-input max)
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mx[c] = sf[c * 32];
#pragma unroll
        for (int i = 1; i < 31; i += 2) mx[c] = fmaxf(mx[c], fmaxf(sf[c * 32 + i], sf[c * 32 + i + 1]));
        mx[c] = fmaxf(mx[c], sf[c * 32 + 31]);
      }
      const float m_blk = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      // P_t(j-1) V has been consumed from smem / accumulated in TMEM before P_t(j) is written or O_t is rescaled
      if (j > 0) {
        mbar_wait(&pv_full[t], (j - 1) & 1);
        tc_fence_after();
      }
      if (__any_sync(0xffffffffu, m_blk > m_run + thresh)) {   // warp-uniform (TMEM accesses are warp-collective)
        const float m_new = fmaxf(m_run, m_blk);
        const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);   // first tile: exp2(-inf) = 0
        l_run *= alpha;
        m_run = m_new;
        if (j > 0) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t o[32];
            tmem_ld_32x32(t_pv + hh * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_pv + hh * 32, o);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = -m_run * p.scale_log2;
      if (kPingPong) {   // wait for the other group to leave its exponential phase (group 1 hands over first)
        if (t == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      // ---- P = exp2(S*scale - m*scale) -> fp16 -> smem (K-major SW128, 2 blocks of 64 kv), row sum in half2 trees.
      // Hand-scheduled with volatile asm (program order is kept): the 32 exponentials of chunk c+1 are issued
      // between the packs of chunk c, so a pack never waits on a MUFU issued just before it (the compiler's own
      // schedule put every F2FP right behind its two MUFUs: one MUFU latency per pair, XU 60 % busy).
      float rs = 0.f;
      auto scale_chunk = [&](int c) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          ffma2(sf[c * 32 + 2 * i], sf[c * 32 + 2 * i + 1], sf[c * 32 + 2 * i], sf[c * 32 + 2 * i + 1], p.scale_log2, neg_m);
      };
      auto exp_pair = [&](int c, int i) {   // in place
        if (kPolyMod > 0 && (i % (kPolyMod > 0 ? kPolyMod : 1)) == kPolyMod - 1) {
          sf[c * 32 + 2 * i] = ex2_poly(sf[c * 32 + 2 * i]);
          sf[c * 32 + 2 * i + 1] = ex2_poly(sf[c * 32 + 2 * i + 1]);
        } else {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(sf[c * 32 + 2 * i]));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(sf[c * 32 + 2 * i + 1]));
        }
      };
      scale_chunk(0);
#pragma unroll
      for (int i = 0; i < 16; ++i) exp_pair(0, i);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < 3) scale_chunk(c + 1);
        uint32_t ph2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c < 3) exp_pair(c + 1, i);
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph2[i]) : "f"(sf[c * 32 + 2 * i + 1]), "f"(sf[c * 32 + 2 * i]));
        }
        uint32_t a0 = hadd2_u32(hadd2_u32(ph2[0], ph2[1]), hadd2_u32(ph2[2], ph2[3]));
        uint32_t a1 = hadd2_u32(hadd2_u32(ph2[4], ph2[5]), hadd2_u32(ph2[6], ph2[7]));
        uint32_t a2 = hadd2_u32(hadd2_u32(ph2[8], ph2[9]), hadd2_u32(ph2[10], ph2[11]));
        uint32_t a3 = hadd2_u32(hadd2_u32(ph2[12], ph2[13]), hadd2_u32(ph2[14], ph2[15]));
        rs += (half2_sum_f32(a0) + half2_sum_f32(a1)) + (half2_sum_f32(a2) + half2_sum_f32(a3));
        const uint32_t blk = p_row + (c >> 1) * kFaTile;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;   // 16 B chunk inside the 128 B row of this block
          st_shared_v4(blk + ((chunk << 4) ^ p_swz), ph2[q * 4 + 0], ph2[q * 4 + 1], ph2[q * 4 + 2], ph2[q * 4 + 3]);
        }
      }
      if (kPingPong) {
        if (t == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
        else if (j + 1 < n) asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
      l_run += rs;
      // ---- publish P_t(j): smem writes visible to the tensor core (async proxy), TMEM accesses retired
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- output: O_t / l from TMEM (each thread writes its 128-byte output row)
    mbar_wait(&pv_full[t], (n - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    uint32_t oa[32], ob[32];
    tmem_ld_32x32(t_pv, oa);
    tmem_ld_32x32(t_pv + 32, ob);
    tmem_ld_wait();
    if (qrow < p.Nq) {
      bf16* dst = p.O + ((long long)b * p.Nq + qrow) * p.ldo + h * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(oa[q * 8 + 0]) * inv, __uint_as_float(oa[q * 8 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(oa[q * 8 + 2]) * inv, __uint_as_float(oa[q * 8 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(oa[q * 8 + 4]) * inv, __uint_as_float(oa[q * 8 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(oa[q * 8 + 6]) * inv, __uint_as_float(oa[q * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[q] = u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(ob[q * 8 + 0]) * inv, __uint_as_float(ob[q * 8 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(ob[q * 8 + 2]) * inv, __uint_as_float(ob[q * 8 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(ob[q * 8 + 4]) * inv, __uint_as_float(ob[q * 8 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(ob[q * 8 + 6]) * inv, __uint_as_float(ob[q * 8 + 7]) * inv);
        reinterpret_cast<uint4*>(dst)[4 + q] = u;
      }
    }
    } else {
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    for (int j = 0; j < n; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const int kv0 = j * 128;
      const bool tail = (kv0 + 128 > p.Nk);
      // ---- pass 1: row max (two 32-column TMEM loads in flight per wait)
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; c += 2) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_s + c * 32, ra);
        tmem_ld_32x32(t_s + c * 32 + 32, rb);
        tmem_ld_wait();
        if (!tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(ra[i]), __uint_as_float(rb[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (kv0 + c * 32 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(ra[i]));
            if (kv0 + c * 32 + 32 + i < p.Nk) mx = fmaxf(mx, __uint_as_float(rb[i]));
          }
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx((m_run - m_new) * p.scale_log2);   // first tile: exp2(-inf) = 0
      const float neg_m = -m_new * p.scale_log2;
      m_run = m_new;
      // ---- fold in the previous tile's P V (its MMA has been running during pass 1)
      if (j > 0) {
        mbar_wait(&pv_full[t], (j - 1) & 1);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_pv, ra);
        tmem_ld_32x32(t_pv + 32, rb);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          o_acc[i] = fmaf(o_acc[i], alpha_prev, __uint_as_float(ra[i]));
          o_acc[32 + i] = fmaf(o_acc[32 + i], alpha_prev, __uint_as_float(rb[i]));
        }
      }
      alpha_prev = alpha;
      // ---- pass 2: P = exp2(S*scale - m*scale) -> fp16 -> smem (K-major SW128, 2 blocks of 64 kv), row sum.
      // Two exponentials per MUFU op (ex2.approx.f16x2); the TMEM load of chunk c+1 is in flight meanwhile.
      float rs = 0.f;
      uint32_t cur[32], nxt[32];
      tmem_ld_32x32(t_s, cur);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < 3) tmem_ld_32x32(t_s + (c + 1) * 32, nxt);
        uint32_t ph2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0 = fmaf(__uint_as_float(cur[2 * i]), p.scale_log2, neg_m);
          float x1 = fmaf(__uint_as_float(cur[2 * i + 1]), p.scale_log2, neg_m);
          if (tail) {
            if (kv0 + c * 32 + 2 * i >= p.Nk) x0 = -INFINITY;
            if (kv0 + c * 32 + 2 * i + 1 >= p.Nk) x1 = -INFINITY;
          }
          ph2[i] = ex2_f16x2(x0, x1);
        }
        // row sum: 4 independent half2 accumulators of 4 pairs each, widened to fp32 per chunk
        uint32_t a0 = hadd2_u32(hadd2_u32(ph2[0], ph2[1]), hadd2_u32(ph2[2], ph2[3]));
        uint32_t a1 = hadd2_u32(hadd2_u32(ph2[4], ph2[5]), hadd2_u32(ph2[6], ph2[7]));
        uint32_t a2 = hadd2_u32(hadd2_u32(ph2[8], ph2[9]), hadd2_u32(ph2[10], ph2[11]));
        uint32_t a3 = hadd2_u32(hadd2_u32(ph2[12], ph2[13]), hadd2_u32(ph2[14], ph2[15]));
        rs += (half2_sum_f32(a0) + half2_sum_f32(a1)) + (half2_sum_f32(a2) + half2_sum_f32(a3));
        uint8_t* blk = p_base + (c >> 1) * kFaTile + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = ph2[q * 4 + 0];
          u.y = ph2[q * 4 + 1];
          u.z = ph2[q * 4 + 2];
          u.w = ph2[q * 4 + 3];
          const int chunk = (c & 1) * 4 + q;   // 16 B chunk inside the 128 B row of this block
          *reinterpret_cast<uint4*>(blk + ((chunk ^ (r & 7)) << 4)) = u;
        }
        if (c < 3) {
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) cur[i] = nxt[i];
        }
        if (c == 2) {   // the last chunk of S_t(j) is now in registers: the tensor core may overwrite S_t
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
        }
      }
      l_run = l_run * alpha + rs;
      // ---- publish P_t(j): smem writes visible to the tensor core (async proxy), TMEM reads retired
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- last tile's P V, normalise, store (each thread writes its 128-byte output row)
    mbar_wait(&pv_full[t], (n - 1) & 1);
    tc_fence_after();
    {
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_pv, ra);
      tmem_ld_32x32(t_pv + 32, rb);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        o_acc[i] = fmaf(o_acc[i], alpha_prev, __uint_as_float(ra[i]));
        o_acc[32 + i] = fmaf(o_acc[32 + i], alpha_prev, __uint_as_float(rb[i]));
      }
    }
    const float inv = 1.f / l_run;
    if (qrow < p.Nq) {
      bf16* dst = p.O + ((long long)b * p.Nq + qrow) * p.ldo + h * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 u;
        u.x = pack_bf16x2(o_acc[q * 8 + 0] * inv, o_acc[q * 8 + 1] * inv);
        u.y = pack_bf16x2(o_acc[q * 8 + 2] * inv, o_acc[q * 8 + 3] * inv);
        u.z = pack_bf16x2(o_acc[q * 8 + 4] * inv, o_acc[q * 8 + 5] * inv);
        u.w = pack_bf16x2(o_acc[q * 8 + 6] * inv, o_acc[q * 8 + 7] * inv);
        reinterpret_cast<uint4*>(dst)[q] = u;
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Host: tensor maps over the strided Q / K / V views (cols, rows-per-batch, batch), box 64 x 128 x 1.
int launch_attention64_tcgen05(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo,
                               int B, int heads, int Nq, int Nk, float scale, cudaStream_t stream) {
  typedef void (*FaKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const FaParams);
  static FaKernel kern = nullptr;
  if (!kern) {
    const char* e1 = getenv("GDF_FA_V1");
    const char* ep = getenv("GDF_FA_POLY");
    const int poly = ep ? atoi(ep) : 0;
    const char* epp = getenv("GDF_FA_PP");
    const bool pp = epp && epp[0] == '1';   // measured slower (612 vs 551 us at N = 4096): off unless asked for
    if (e1 && e1[0] == '1') kern = attention64_tcgen05_kernel<1, 0, false>;
    else if (poly == 4) kern = pp ? attention64_tcgen05_kernel<2, 4, true> : attention64_tcgen05_kernel<2, 4, false>;
    else if (pp) kern = attention64_tcgen05_kernel<2, 0, true>;
    else kern = attention64_tcgen05_kernel<2, 0, false>;
    GDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem));
  }
  CUtensorMap mq, mk, mv;
  const uint32_t box[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)Nq, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)Nq * ldq * 2};
    GDF_TRY(make_tmap_bf16(&mq, Q, 3, dims, str, box));
  }
  {
    uint64_t dims[3] = {(uint64_t)heads * 64, (uint64_t)Nk, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)ldk * 2, (uint64_t)Nk * ldk * 2};
    GDF_TRY(make_tmap_bf16(&mk, K, 3, dims, str, box));
    uint64_t strv[2] = {(uint64_t)ldv * 2, (uint64_t)Nk * ldv * 2};
    GDF_TRY(make_tmap_bf16(&mv, V, 3, dims, strv, box));
  }
  FaParams p;
  p.Nq = Nq;
  p.Nk = Nk;
  p.heads = heads;
  p.num_kv_tiles = (Nk + 127) / 128;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.O = O;
  p.ldo = ldo;
  dim3 grid((Nq + 255) / 256, heads, B);
  GDF_CUDA(launch_pdl(kern, grid, dim3(kFaThreads), (size_t)kFaSmem, stream, mq, mk, mv, p));
  return GDF_OK;
}

}  // namespace gdf
