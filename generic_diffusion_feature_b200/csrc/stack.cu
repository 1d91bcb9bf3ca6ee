// Feature-stack construction: bilinear resize (align_corners=False, PyTorch semantics) of captured fp16 maps
// + channel concat, in NHWC (pixel-major, feeds the similarity GEMM) and/or NCHW (the reference's layout).
// Reference: F.interpolate(f, (128,128), mode='bilinear') + torch.cat in
// correspondence/correspondence/aggregation_network.py:62-66. HBM-bound; NCHW output is transposed through
// shared memory so both the reads (channel-contiguous) and the writes (pixel-contiguous) are coalesced.
#include <stdlib.h>
#include "ops.h"

namespace gdf {

struct BilinearTap {
  int i0, i1;
  float w0, w1;
};
// PyTorch area_pixel_compute_source_index for align_corners=False: src = max(0, (dst + 0.5) * scale - 0.5)
__device__ __forceinline__ BilinearTap bilinear_tap(int dst, int in_size, float scale) {
  float src = ((float)dst + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  BilinearTap t;
  t.i0 = (int)src;
  if (t.i0 > in_size - 1) t.i0 = in_size - 1;
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.w1 = src - (float)t.i0;
  t.w0 = 1.f - t.w1;
  return t;
}

// same evaluation order as ATen's upsample_bilinear2d: wy0*(wx0*a + wx1*b) + wy1*(wx0*c + wx1*d)
__device__ __forceinline__ void blend8(const uint4& a, const uint4& b, const uint4& c, const uint4& d, float wy0,
                                       float wy1, float wx0, float wx1, float (&out)[8]) {
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  const __half2* pd = reinterpret_cast<const __half2*>(&d);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = __half22float2(pa[j]), fb = __half22float2(pb[j]);
    const float2 fc = __half22float2(pc[j]), fd = __half22float2(pd[j]);
    out[2 * j] = wy0 * (wx0 * fa.x + wx1 * fb.x) + wy1 * (wx0 * fc.x + wx1 * fd.x);
    out[2 * j + 1] = wy0 * (wx0 * fa.y + wx1 * fb.y) + wy1 * (wx0 * fc.y + wx1 * fd.y);
  }
}

// NHWC -> NHWC: one thread per (output pixel, 8 channels).
__global__ void __launch_bounds__(256)
resize_nhwc_kernel(const __half* __restrict__ src, int h, int w, int C, int c_off, int B, int OH, int OW, int Ctot,
                   __half* __restrict__ out) {
  const int C8 = C / 8;
  const long long total = (long long)B * OH * OW * C8;
  const float sy = (float)h / (float)OH, sx = (float)w / (float)OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const BilinearTap ty = bilinear_tap(oy, h, sy), tx = bilinear_tap(ox, w, sx);
    const uint4* s = reinterpret_cast<const uint4*>(src + ((long long)b * h * w) * C) + c8;
    const uint4 v00 = __ldg(s + ((long long)ty.i0 * w + tx.i0) * C8);
    const uint4 v01 = __ldg(s + ((long long)ty.i0 * w + tx.i1) * C8);
    const uint4 v10 = __ldg(s + ((long long)ty.i1 * w + tx.i0) * C8);
    const uint4 v11 = __ldg(s + ((long long)ty.i1 * w + tx.i1) * C8);
    float o[8];
    blend8(v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1, o);
    uint4 u;
    u.x = pack_f16x2(o[0], o[1]);
    u.y = pack_f16x2(o[2], o[3]);
    u.z = pack_f16x2(o[4], o[5]);
    u.w = pack_f16x2(o[6], o[7]);
    *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * Ctot + c_off + c8 * 8) = u;
  }
}

// NHWC -> NHWC for integer up-scaling factors S = OH / h = OW / w (1, 2, 4, 8: every map of the SDXL / SD-2.1 stacks):
// one thread per (source cell, 8 channels). All output pixels whose bilinear taps are the four corners of that cell -
// oy in [S*cy + S/2, S*cy + 3S/2), cy = -1 .. h-1 with clamped corners - are produced from ONE load of the corners:
// the general kernel above re-reads them S*S times through L2 (4x16 B read per 16 B written), which capped it at a
// quarter of the HBM write rate. Weights come from the same bilinear_tap arithmetic, so results are bit-identical.
// The warp's lanes hold consecutive channel groups of one cell (the channel index space is padded to a multiple of 32),
// so stores are 512 B contiguous per warp and the per-pixel squared norm (sumsq, optional) is a shuffle reduction +
// one atomicAdd per (warp, pixel) instead of a second pass over the stack.
template <int S>
__global__ void __launch_bounds__(256)
resize_cell_nhwc_kernel(const __half* __restrict__ src, int h, int w, int C, int c_off, int B, int Ctot,
                        __half* __restrict__ out, float* __restrict__ sumsq) {
  const int OH = h * S, OW = w * S;
  const int C8 = C / 8, C8P = (C8 + 31) & ~31;
  const int ch = (S == 1) ? h : h + 1, cw = (S == 1) ? w : w + 1;     // cells per axis
  // 32-bit index arithmetic (the host checks total < 2^31): the four 64-bit divisions per cell of the first version
  // were about as many instructions as the blends of an S = 2 cell
  const unsigned total = (unsigned)B * ch * cw * C8P;
  const int lane = threadIdx.x & 31;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned q = i / (unsigned)C8P;
    const int c8x = (int)(i - q * (unsigned)C8P);
    unsigned r = q;
    const int cxi = (int)(r % (unsigned)cw);
    r /= (unsigned)cw;
    const int cyi = (int)(r % (unsigned)ch);
    const int b = (int)(r / (unsigned)ch);
    const bool live = c8x < C8;
    const int cy = (S == 1) ? cyi : cyi - 1, cx = (S == 1) ? cxi : cxi - 1;
    const int r0 = cy < 0 ? 0 : cy, r1 = cy + 1 > h - 1 ? h - 1 : cy + 1;
    const int q0 = cx < 0 ? 0 : cx, q1 = cx + 1 > w - 1 ? w - 1 : cx + 1;
    uint4 v00 = make_uint4(0, 0, 0, 0), v01 = v00, v10 = v00, v11 = v00;
    if (live) {
      const uint4* sp = reinterpret_cast<const uint4*>(src + ((long long)b * h * w) * C) + c8x;
      v00 = __ldg(sp + ((long long)r0 * w + q0) * C8);
      if (S > 1) {
        v01 = __ldg(sp + ((long long)r0 * w + q1) * C8);
        v10 = __ldg(sp + ((long long)r1 * w + q0) * C8);
        v11 = __ldg(sp + ((long long)r1 * w + q1) * C8);
      }
    }
    const int oy0 = (S == 1) ? cy : S * cy + S / 2, ox0 = (S == 1) ? cx : S * cx + S / 2;
    float txw0[S], txw1[S];                        // x taps of the S output columns of this cell (same for every dy)
#pragma unroll
    for (int dx = 0; dx < S; ++dx) {
      const BilinearTap tx = bilinear_tap(ox0 + dx < 0 ? 0 : ox0 + dx, w, 1.f / (float)S);
      txw0[dx] = tx.w0;
      txw1[dx] = tx.w1;
    }
#pragma unroll
    for (int dy = 0; dy < S; ++dy) {
      const int oy = oy0 + dy;
      if (oy < 0 || oy >= OH) continue;            // warp-uniform (one cell per warp)
      const BilinearTap ty = bilinear_tap(oy, h, 1.f / (float)S);
#pragma unroll
      for (int dx = 0; dx < S; ++dx) {
        const int ox = ox0 + dx;
        if (ox < 0 || ox >= OW) continue;
        float ss = 0.f;
        if (live) {
          uint4 u;
          if (S == 1) {
            u = v00;
            if (sumsq) {
              const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(hp[j]);
                ss += f.x * f.x + f.y * f.y;
              }
            }
          } else {
            float o[8];
            blend8(v00, v01, v10, v11, ty.w0, ty.w1, txw0[dx], txw1[dx], o);
            u.x = pack_f16x2(o[0], o[1]);
            u.y = pack_f16x2(o[2], o[3]);
            u.z = pack_f16x2(o[4], o[5]);
            u.w = pack_f16x2(o[6], o[7]);
            if (sumsq) {   // norm of the ROUNDED values (what a second pass over the fp16 stack would read)
              const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(hp[j]);
                ss += f.x * f.x + f.y * f.y;
              }
            }
          }
          *reinterpret_cast<uint4*>(out + (((long long)b * OH + oy) * OW + ox) * Ctot + c_off + c8x * 8) = u;
        }
        if (sumsq) {
#pragma unroll
          for (int o2 = 16; o2 > 0; o2 >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o2);
          if (lane == 0) atomicAdd(sumsq + ((long long)b * OH + oy) * OW + ox, ss);
        }
      }
    }
  }
}

// NHWC -> NHWC, EVERY map of the stack in one launch: block = 8 x 8 output pixels x all channels. The per-map kernels
// above write one map's channel slab of every pixel (512 B pieces 2 * Ctot bytes apart: 2.5-2.9 TB/s, the same as a
// strided torch copy into the stack); here a block produces the whole 2 * Ctot byte row of its 64 pixels within a few
// microseconds, so DRAM sees full-row writes. Warp w owns tile row w: for every map, for each of its 8 pixels the taps
// are computed once (warp-uniform) and the lanes walk the map's channel vectors (32 x 16 B = 512 B per step); the four
// corners of neighbouring output pixels are the same source pixels (scale factors 1 ... 8), so the corner loads hit L1
// after the first touch. The per-pixel squared norm is a register sum per (warp, pixel) + one shuffle reduction at the
// end: no atomics, no memset, no second pass. Any scale factor (general taps), C and c_off multiples of 8.
constexpr int kMaxFusedSrc = 64;
struct ResizeSrcDev {
  const __half* ptr;
  int h, w, C, c_off;
  float sy, sx;
};
struct ResizeSrcList {
  ResizeSrcDev s[kMaxFusedSrc];
  int n;
};
__global__ void __launch_bounds__(256)
resize_concat_tile_kernel(const __grid_constant__ ResizeSrcList L, int B, int OH, int OW, int Ctot,
                          __half* __restrict__ out, float* __restrict__ sumsq, int sumsq_accumulate) {
  const int tiles_x = (OW + 7) >> 3, tiles_y = (OH + 7) >> 3;
  const int tx_ = blockIdx.x % tiles_x;
  const int ty_ = (blockIdx.x / tiles_x) % tiles_y;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oy = ty_ * 8 + warp;
  if (oy >= OH) return;
  const int ox0 = tx_ * 8;
  float ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ss[j] = 0.f;
  __half* orow = out + (((long long)b * OH + oy) * OW + ox0) * Ctot;
  for (int m = 0; m < L.n; ++m) {
    const ResizeSrcDev& s = L.s[m];
    const int C8 = s.C >> 3;
    const BilinearTap ty = bilinear_tap(oy, s.h, s.sy);
    const uint4* base = reinterpret_cast<const uint4*>(s.ptr + ((long long)b * s.h * s.w) * s.C);
    const uint4* r0 = base + (long long)ty.i0 * s.w * C8;
    const uint4* r1 = base + (long long)ty.i1 * s.w * C8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (ox0 + j >= OW) break;
      const BilinearTap tx = bilinear_tap(ox0 + j, s.w, s.sx);
      const uint4* p00 = r0 + (long long)tx.i0 * C8;
      const uint4* p01 = r0 + (long long)tx.i1 * C8;
      const uint4* p10 = r1 + (long long)tx.i0 * C8;
      const uint4* p11 = r1 + (long long)tx.i1 * C8;
      uint4* dst = reinterpret_cast<uint4*>(orow + (long long)j * Ctot + s.c_off);
      for (int c8 = lane; c8 < C8; c8 += 32) {
        const uint4 v00 = __ldg(p00 + c8), v01 = __ldg(p01 + c8), v10 = __ldg(p10 + c8), v11 = __ldg(p11 + c8);
        float o[8];
        blend8(v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1, o);
        uint4 u;
        u.x = pack_f16x2(o[0], o[1]);
        u.y = pack_f16x2(o[2], o[3]);
        u.z = pack_f16x2(o[4], o[5]);
        u.w = pack_f16x2(o[6], o[7]);
        if (sumsq) {   // norm of the ROUNDED values (what a second pass over the fp16 stack would read)
          const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(hp[q]);
            ss[j] += f.x * f.x + f.y * f.y;
          }
        }
        dst[c8] = u;
      }
    }
  }
  if (sumsq) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = ss[j];
#pragma unroll
      for (int o2 = 16; o2 > 0; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
      if (lane == 0 && ox0 + j < OW) {   // stacks of more than kMaxFusedSrc maps: later launches add their channels' share
        float* d = sumsq + ((long long)b * OH + oy) * OW + ox0 + j;
        *d = sumsq_accumulate ? *d + v : v;
      }
    }
  }
}

// any channel count / offset (latent-sized maps: unet-in / unet-out have 4 channels): one thread per output element
__global__ void __launch_bounds__(256)
resize_nhwc_scalar_kernel(const __half* __restrict__ src, int h, int w, int C, int c_off, int B, int OH, int OW, int Ctot,
                          __half* __restrict__ out) {
  const long long total = (long long)B * OH * OW * C;
  const float sy = (float)h / (float)OH, sx = (float)w / (float)OW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const BilinearTap ty = bilinear_tap(oy, h, sy), tx = bilinear_tap(ox, w, sx);
    const __half* s = src + ((long long)b * h * w) * C + c;
    const float a = __half2float(s[((long long)ty.i0 * w + tx.i0) * C]), bq = __half2float(s[((long long)ty.i0 * w + tx.i1) * C]);
    const float cq = __half2float(s[((long long)ty.i1 * w + tx.i0) * C]), d = __half2float(s[((long long)ty.i1 * w + tx.i1) * C]);
    out[(((long long)b * OH + oy) * OW + ox) * Ctot + c_off + c] =
        __float2half_rn(ty.w0 * (tx.w0 * a + tx.w1 * bq) + ty.w1 * (tx.w0 * cq + tx.w1 * d));
  }
}

template <int S>
static bool launch_resize_cell(const ResizeSrc& s, int B, int Ctot, __half* out, float* sumsq, cudaStream_t stream) {
  const int C8P = (s.C / 8 + 31) & ~31;
  const long long total = (long long)B * (S == 1 ? s.h : s.h + 1) * (S == 1 ? s.w : s.w + 1) * C8P;
  if (total >= (1ll << 31)) return false;          // the kernel indexes cells with 32 bits
  const long long blocks = (total + 255) / 256;
  resize_cell_nhwc_kernel<S><<<(unsigned)(blocks < 148 * 64 ? blocks : 148 * 64), 256, 0, stream>>>(
      s.ptr, s.h, s.w, s.C, s.c_off, B, Ctot, out, sumsq);
  return true;
}

// NHWC -> NCHW through shared memory: block = 64 output pixels (one row segment) x 64 channels. Phase 1 blends
// (pixel, 8 channels) items and scatters them into a [channel][pixel] tile whose 8-pixel groups are XOR-swizzled by the
// channel vector index (conflict-free 2-byte scatter); phase 2 writes 8 pixels of one channel per thread as one 16 B
// store: a warp covers 4 channels x 128 B contiguous (the first version wrote 64 B per warp row with 2-byte stores and
// ran at 18 % of the HBM peak).
__global__ void __launch_bounds__(256)
resize_nchw_kernel(const __half* __restrict__ src, int h, int w, int C, int c_off, int B, int OH, int OW, int Ctot,
                   __half* __restrict__ out) {
  __shared__ __align__(16) __half tile[64][64];
  const int xblocks = (OW + 63) / 64;
  const int xb = blockIdx.x % xblocks;
  const int oy = (blockIdx.x / xblocks) % OH;
  const int b = blockIdx.x / (xblocks * OH);
  const int cb = blockIdx.y * 64;
  const float sy = (float)h / (float)OH, sx = (float)w / (float)OW;
  const int C8 = C / 8;
  const BilinearTap ty = bilinear_tap(oy, h, sy);
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // 64 pixels x 8 channel-vectors = 512 work items
    const int item = threadIdx.x + k * 256;
    const int px = item >> 3, cv = item & 7;
    const int ox = xb * 64 + px;
    const int c8 = cb / 8 + cv;
    if (ox < OW && c8 < C8) {
      const BilinearTap tx = bilinear_tap(ox, w, sx);
      const uint4* s = reinterpret_cast<const uint4*>(src + ((long long)b * h * w) * C) + c8;
      const uint4 v00 = __ldg(s + ((long long)ty.i0 * w + tx.i0) * C8);
      const uint4 v01 = __ldg(s + ((long long)ty.i0 * w + tx.i1) * C8);
      const uint4 v10 = __ldg(s + ((long long)ty.i1 * w + tx.i0) * C8);
      const uint4 v11 = __ldg(s + ((long long)ty.i1 * w + tx.i1) * C8);
      float o[8];
      blend8(v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1, o);
      const int col = (((px >> 3) ^ cv) << 3) | (px & 7);
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[cv * 8 + j][col] = __float2half_rn(o[j]);
    }
  }
  __syncthreads();
  const bool vec = (OW % 8 == 0);
#pragma unroll
  for (int k = 0; k < 2; ++k) {  // 64 channels x 8 pixel groups, pixel group fastest
    const int item = threadIdx.x + k * 256;
    const int ch = item >> 3, pg = item & 7;
    const int ox = xb * 64 + pg * 8;
    if (cb + ch >= C || ox >= OW) continue;
    const __half* t = &tile[ch][(pg ^ (ch >> 3)) << 3];
    __half* dst = out + (((long long)b * Ctot + c_off + cb + ch) * OH + oy) * OW + ox;
    if (vec && ox + 8 <= OW) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(t);
    } else {
      for (int j = 0; j < 8 && ox + j < OW; ++j) dst[j] = t[j];
    }
  }
}

// per-pixel squared L2 norm of an NHWC fp16 stack: one warp per pixel.
__global__ void __launch_bounds__(256)
rownorm_kernel(const __half* __restrict__ x, long long rows, int C, float* __restrict__ sumsq) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* p = reinterpret_cast<const uint4*>(x + row * C);
  float s = 0.f;
  for (int v = lane; v < C / 8; v += 32) {
    const uint4 u = __ldg(p + v);
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(hp[j]);
      s += f.x * f.x + f.y * f.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) sumsq[row] = s;
}

cudaError_t launch_resize_concat(const ResizeSrc* srcs_host, int n_src, int B, int OH, int OW, int Ctot,
                                 __half* out_nhwc, __half* out_nchw, float* sumsq, cudaStream_t stream) {
  static int use_cell = -1;
  if (use_cell < 0) {
    const char* e = getenv("GDF_RESIZE_CELL");     // 0: the general kernel for every map (A/B timing)
    use_cell = e ? atoi(e) : 1;
  }
  // per-pixel squared norms: accumulated by the cell kernels (atomicAdd) when every map takes that path, otherwise
  // by one pass over the finished stack (rownorm_kernel)
  bool all_cell = use_cell != 0 && out_nhwc != nullptr;
  for (int i = 0; i < n_src && all_cell; ++i) {
    const ResizeSrc& s = srcs_host[i];
    const int sc = (s.h > 0 && OH % s.h == 0) ? OH / s.h : 0;
    all_cell = (OW == s.w * sc) && (sc == 1 || sc == 2 || sc == 4 || sc == 8) && s.C % 8 == 0 && s.c_off % 8 == 0 &&
               Ctot % 8 == 0 && (long long)B * (s.h + 1) * (s.w + 1) * ((s.C / 8 + 31) & ~31) < (1ll << 31);
  }
  // every map vectorisable and at most kMaxFusedSrc of them: ONE launch for the NHWC stack (+ sumsq); GDF_RESIZE_FUSED=0
  // keeps the per-map kernels (A/B timing)
  static int use_fused = -1;
  if (use_fused < 0) {
    const char* e = getenv("GDF_RESIZE_FUSED");
    use_fused = e ? atoi(e) : 1;
  }
  bool fused = use_fused != 0 && out_nhwc != nullptr && Ctot % 8 == 0;
  for (int i = 0; i < n_src && fused; ++i)
    fused = srcs_host[i].C % 8 == 0 && srcs_host[i].c_off % 8 == 0 && srcs_host[i].h > 0 && srcs_host[i].w > 0;
  if (fused) {
    const unsigned blocks = (unsigned)(((OW + 7) / 8) * ((OH + 7) / 8) * B);
    for (int i0 = 0; i0 < n_src; i0 += kMaxFusedSrc) {   // kMaxFusedSrc maps per launch (kernel-parameter space)
      ResizeSrcList L;
      L.n = n_src - i0 < kMaxFusedSrc ? n_src - i0 : kMaxFusedSrc;
      for (int i = 0; i < L.n; ++i) {
        const ResizeSrc& s = srcs_host[i0 + i];
        L.s[i].ptr = s.ptr;
        L.s[i].h = s.h;
        L.s[i].w = s.w;
        L.s[i].C = s.C;
        L.s[i].c_off = s.c_off;
        L.s[i].sy = (float)s.h / (float)OH;
        L.s[i].sx = (float)s.w / (float)OW;
      }
      resize_concat_tile_kernel<<<blocks, 256, 0, stream>>>(L, B, OH, OW, Ctot, out_nhwc, sumsq, i0 > 0 ? 1 : 0);
    }
    out_nhwc = nullptr;    // done (sumsq too); the reference-layout output, if requested, follows below
    sumsq = nullptr;
  }
  float* fused_sumsq = (sumsq && all_cell) ? sumsq : nullptr;
  if (fused_sumsq) {
    cudaError_t e = cudaMemsetAsync(sumsq, 0, (size_t)B * OH * OW * sizeof(float), stream);
    if (e != cudaSuccess) return e;
  }
  for (int i = 0; i < n_src; ++i) {
    const ResizeSrc& s = srcs_host[i];
    if (s.C % 8 != 0 || s.c_off % 8 != 0 || Ctot % 8 != 0) {
      // maps that cannot use 128-bit channel vectors (4-channel latents, or anything stacked behind them)
      if (out_nchw) return cudaErrorInvalidValue;     // the reference-layout transpose kernel is vector-only
      if (out_nhwc) {
        const long long total = (long long)B * OH * OW * s.C;
        const long long blocks = (total + 255) / 256;
        resize_nhwc_scalar_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, stream>>>(
            s.ptr, s.h, s.w, s.C, s.c_off, B, OH, OW, Ctot, out_nhwc);
      }
      continue;
    }
    if (out_nhwc) {
      const int sc = (use_cell && s.h > 0 && OH % s.h == 0 && OW == s.w * (OH / s.h)) ? OH / s.h : 0;
      bool done = false;
      if (sc == 1) done = launch_resize_cell<1>(s, B, Ctot, out_nhwc, fused_sumsq, stream);
      else if (sc == 2) done = launch_resize_cell<2>(s, B, Ctot, out_nhwc, fused_sumsq, stream);
      else if (sc == 4) done = launch_resize_cell<4>(s, B, Ctot, out_nhwc, fused_sumsq, stream);
      else if (sc == 8) done = launch_resize_cell<8>(s, B, Ctot, out_nhwc, fused_sumsq, stream);
      if (!done) {
        if (fused_sumsq) return cudaErrorInvalidValue;   // (cannot happen: all_cell implies 32-bit cell counts below)
        const long long total = (long long)B * OH * OW * (s.C / 8);
        const long long blocks = (total + 255) / 256;
        resize_nhwc_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, stream>>>(
            s.ptr, s.h, s.w, s.C, s.c_off, B, OH, OW, Ctot, out_nhwc);
      }
    }
    if (out_nchw) {
      dim3 grid((unsigned)(((OW + 63) / 64) * OH * B), (unsigned)((s.C + 63) / 64));
      resize_nchw_kernel<<<grid, 256, 0, stream>>>(s.ptr, s.h, s.w, s.C, s.c_off, B, OH, OW, Ctot, out_nchw);
    }
  }
  if (sumsq && out_nhwc && !fused_sumsq) {
    const long long rows = (long long)B * OH * OW;
    rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(out_nhwc, rows, Ctot, sumsq);
  }
  return cudaGetLastError();
}

}  // namespace gdf

// =============================================================================================== correspondence
// find_nn_source_correspondences (correspondence/correspondence/correspondence_utils.py:113-138) upsamples both
// stacks to load x load, gathers the query rows, L2-normalises, multiplies (n x load^2 x C GEMM) and takes the
// arg-max. Bilinear interpolation is linear, so the same similarities are obtained from the LOW-resolution product
// q . F2^T (n x hw^2 x C, 16x fewer FLOPs at 128 -> 512) interpolated per query, divided by the norm of the
// interpolated target vector, which follows from 5 Gram maps of neighbouring low-resolution pixels.
namespace gdf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// q[i, :] = bilinear_up(stack1)[query_i] as fp16 (one warp per query)
__global__ void __launch_bounds__(256)
corr_gather_queries_kernel(const __half* __restrict__ s1, int C, int hw, int load, const int* __restrict__ qyx, int n,
                           __half* __restrict__ q) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float sc = (float)hw / (float)load;
  const BilinearTap ty = bilinear_tap(qyx[2 * i], hw, sc), tx = bilinear_tap(qyx[2 * i + 1], hw, sc);
  const int C8 = C / 8;
  const uint4* b = reinterpret_cast<const uint4*>(s1);
  for (int v = lane; v < C8; v += 32) {
    const uint4 v00 = __ldg(b + ((long long)ty.i0 * hw + tx.i0) * C8 + v);
    const uint4 v01 = __ldg(b + ((long long)ty.i0 * hw + tx.i1) * C8 + v);
    const uint4 v10 = __ldg(b + ((long long)ty.i1 * hw + tx.i0) * C8 + v);
    const uint4 v11 = __ldg(b + ((long long)ty.i1 * hw + tx.i1) * C8 + v);
    float o[8];
    blend8(v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1, o);
    uint4 u;
    u.x = pack_f16x2(o[0], o[1]);
    u.y = pack_f16x2(o[2], o[3]);
    u.z = pack_f16x2(o[4], o[5]);
    u.w = pack_f16x2(o[6], o[7]);
    reinterpret_cast<uint4*>(q + (long long)i * C)[v] = u;
  }
}

// Gram maps of stack2: g[0] = <f,f>, g[1] = <f, right>, g[2] = <f, down>, g[3] = <f, down-right>,
// g[4] = <right, down> (one warp per low-resolution pixel; neighbours clamped at the border)
__global__ void __launch_bounds__(256)
corr_gram_kernel(const __half* __restrict__ s2, int C, int hw, float* __restrict__ gram) {
  const int lane = threadIdx.x & 31;
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (p >= hw * hw) return;
  const int y = p / hw, x = p % hw;
  const int x1 = min(x + 1, hw - 1), y1 = min(y + 1, hw - 1);
  const int C8 = C / 8;
  const uint4* b = reinterpret_cast<const uint4*>(s2);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
  for (int v = lane; v < C8; v += 32) {
    const uint4 f = __ldg(b + ((long long)y * hw + x) * C8 + v);
    const uint4 r = __ldg(b + ((long long)y * hw + x1) * C8 + v);
    const uint4 d = __ldg(b + ((long long)y1 * hw + x) * C8 + v);
    const uint4 e = __ldg(b + ((long long)y1 * hw + x1) * C8 + v);
    const __half2* pf = reinterpret_cast<const __half2*>(&f);
    const __half2* pr = reinterpret_cast<const __half2*>(&r);
    const __half2* pd = reinterpret_cast<const __half2*>(&d);
    const __half2* pe = reinterpret_cast<const __half2*>(&e);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 ff = __half22float2(pf[j]), fr = __half22float2(pr[j]);
      const float2 fd = __half22float2(pd[j]), fe = __half22float2(pe[j]);
      a0 += ff.x * ff.x + ff.y * ff.y;
      a1 += ff.x * fr.x + ff.y * fr.y;
      a2 += ff.x * fd.x + ff.y * fd.y;
      a3 += ff.x * fe.x + ff.y * fe.y;
      a4 += fr.x * fd.x + fr.y * fd.y;
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
  if (lane == 0) {
    const int n = hw * hw;
    gram[p] = a0;
    gram[n + p] = a1;
    gram[2 * n + p] = a2;
    gram[3 * n + p] = a3;
    gram[4 * n + p] = a4;
  }
}

// inverse norm of the bilinearly interpolated target vector at every load x load position
__global__ void __launch_bounds__(256)
corr_invnorm_kernel(const float* __restrict__ gram, int hw, int load, float* __restrict__ invnorm) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= load * load) return;
  const int oy = p / load, ox = p % load;
  const float sc = (float)hw / (float)load;
  const BilinearTap ty = bilinear_tap(oy, hw, sc), tx = bilinear_tap(ox, hw, sc);
  const int n = hw * hw;
  const float* G0 = gram;
  const float* Gx = gram + n;
  const float* Gy = gram + 2 * n;
  const float* Gd = gram + 3 * n;
  const float* Ga = gram + 4 * n;
  const int i00 = ty.i0 * hw + tx.i0, i01 = ty.i0 * hw + tx.i1, i10 = ty.i1 * hw + tx.i0, i11 = ty.i1 * hw + tx.i1;
  const bool dx = tx.i1 != tx.i0, dy = ty.i1 != ty.i0;
  const float a00 = ty.w0 * tx.w0, a01 = ty.w0 * tx.w1, a10 = ty.w1 * tx.w0, a11 = ty.w1 * tx.w1;
  // pairwise inner products of the four taps (coinciding taps at the border fall back to the self product)
  const float g0001 = dx ? Gx[i00] : G0[i00];
  const float g1011 = dx ? Gx[i10] : G0[i10];
  const float g0010 = dy ? Gy[i00] : G0[i00];
  const float g0111 = dy ? Gy[i01] : G0[i01];
  const float g0011 = (dx && dy) ? Gd[i00] : (dx ? Gx[i00] : (dy ? Gy[i00] : G0[i00]));
  const float g0110 = (dx && dy) ? Ga[i00] : (dx ? Gx[i00] : (dy ? Gy[i00] : G0[i00]));
  float s = a00 * a00 * G0[i00] + a01 * a01 * G0[i01] + a10 * a10 * G0[i10] + a11 * a11 * G0[i11];
  s += 2.f * (a00 * a01 * g0001 + a10 * a11 * g1011 + a00 * a10 * g0010 + a01 * a11 * g0111 + a00 * a11 * g0011 +
              a01 * a10 * g0110);
  invnorm[p] = rsqrtf(fmaxf(s, 1e-30f));
}

// arg-max over the load x load grid of bilinear(sims_lr[i]) * invnorm; one block per query, first index wins ties
__global__ void __launch_bounds__(256)
corr_argmax_kernel(const float* __restrict__ sims, const float* __restrict__ invnorm, int hw, int load, int n,
                   long long* __restrict__ idx_out) {
  extern __shared__ float srow[];   // hw*hw low-resolution similarities of this query
  __shared__ float red_v[8];
  __shared__ int red_i[8];
  const int i = blockIdx.x;
  const int nlr = hw * hw;
  for (int k = threadIdx.x; k < nlr; k += blockDim.x) srow[k] = sims[(long long)i * nlr + k];
  __syncthreads();
  const float sc = (float)hw / (float)load;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int row = threadIdx.x >> 5; row < load; row += 8) {
    const BilinearTap ty = bilinear_tap(row, hw, sc);
    const float* r0 = srow + ty.i0 * hw;
    const float* r1 = srow + ty.i1 * hw;
    for (int ox = threadIdx.x & 31; ox < load; ox += 32) {
      const BilinearTap tx = bilinear_tap(ox, hw, sc);
      const float v = (ty.w0 * (tx.w0 * r0[tx.i0] + tx.w1 * r0[tx.i1]) + ty.w1 * (tx.w0 * r1[tx.i0] + tx.w1 * r1[tx.i1])) *
                      __ldg(invnorm + row * load + ox);
      const int p = row * load + ox;
      if (v > best || (v == best && p < best_i)) {
        best = v;
        best_i = p;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    red_v[threadIdx.x >> 5] = best;
    red_i[threadIdx.x >> 5] = best_i;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (red_v[w] > best || (red_v[w] == best && red_i[w] < best_i)) {
        best = red_v[w];
        best_i = red_i[w];
      }
    idx_out[i] = best_i;
  }
}

size_t corr_workspace_floats(int n, int hw, int C) {
  // q (fp16, n*C halves) | sims (n*hw*hw) | gram (5*hw*hw) | invnorm: sized by the caller's load (<= 1024^2)
  return (size_t)n * C / 2 + 64 + (size_t)n * hw * hw + 5 * (size_t)hw * hw + (size_t)1024 * 1024;
}

int launch_correspond(const __half* stack1, const __half* stack2, int C, int hw, int load, const int* query_yx, int n,
                      long long* idx_out, float* workspace, cudaStream_t stream) {
  if (C % 8 != 0 || load > 1024 || hw * hw * 4 > 200 * 1024) return fail(GDF_ERR_SHAPE, "launch_correspond: bad shape");
  __half* q = reinterpret_cast<__half*>(workspace);
  float* sims = workspace + ((size_t)n * C / 2 + 63) / 64 * 64;
  float* gram = sims + (size_t)n * hw * hw;
  float* invnorm = gram + 5 * (size_t)hw * hw;
  corr_gather_queries_kernel<<<(n + 7) / 8, 256, 0, stream>>>(stack1, C, hw, load, query_yx, n, q);
  corr_gram_kernel<<<(hw * hw + 7) / 8, 256, 0, stream>>>(stack2, C, hw, gram);
  corr_invnorm_kernel<<<(load * load + 255) / 256, 256, 0, stream>>>(gram, hw, load, invnorm);
  // similarity GEMM on the tensor cores: sims[n, hw*hw] = q[n, C] . stack2[hw*hw, C]^T (fp16 operands, fp32 out)
  GemmLaunch g;
  Epilogue e;
  e.in_f16 = true;
  e.out_f32 = sims;
  e.ld_out_f32 = hw * hw;
  GDF_TRY(build_linear(&g, reinterpret_cast<const bf16*>(q), n, C, C, reinterpret_cast<const bf16*>(stack2), hw * hw,
                       C, e));
  GDF_CUDA(launch_gemm(g, stream));
  static bool attr = false;
  if (!attr) {
    GDF_CUDA(cudaFuncSetAttribute(corr_argmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  corr_argmax_kernel<<<n, 256, (size_t)hw * hw * 4, stream>>>(sims, invnorm, hw, load, n, idx_out);
  GDF_CUDA(cudaGetLastError());
  return GDF_OK;
}

// ---------------------------------------------------------------- feature_resize (FeatureStore.store, feature_extractor.py:
// 51-53): F.adaptive_avg_pool2d(feat, (h // r, w // r)) on a captured fp16 NHWC map. Window of output cell i along an
// axis of length H: [floor(i * H / OH), ceil((i + 1) * H / OH)) (ATen adaptive pooling). 8 channels per thread, fp32 sums.
__global__ void adaptive_avgpool_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H, int W,
                                             int C, int OH, int OW) {
  const int cv = C >> 3;
  const long long total = (long long)B * OH * OW * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) << 3;
    long long r = i / cv;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH;
    const int x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int yy = y0; yy < y1; ++yy) {
      for (int xx = x0; xx < x1; ++xx) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)b * H + yy) * W + xx) * C + c);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __half22float2(h2[q]);
          acc[2 * q] += f.x;
          acc[2 * q + 1] += f.y;
        }
      }
    }
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) oh[q] = __floats2half2_rn(acc[2 * q] * inv, acc[2 * q + 1] * inv);
    *reinterpret_cast<uint4*>(y + (((long long)b * OH + oy) * OW + ox) * C + c) = o;
  }
}
// same, one channel per thread (maps whose channel count is not a multiple of 8: unet-in / unet-out, 4 channels)
__global__ void adaptive_avgpool_nhwc_scalar_kernel(const __half* __restrict__ x, __half* __restrict__ y, int B, int H,
                                                    int W, int C, int OH, int OW) {
  const long long total = (long long)B * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH;
    const int x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
    float acc = 0.f;
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) acc += __half2float(x[(((long long)b * H + yy) * W + xx) * C + c]);
    y[i] = __float2half_rn(acc / (float)((y1 - y0) * (x1 - x0)));
  }
}
cudaError_t launch_adaptive_avgpool_nhwc(const __half* x, __half* y, int B, int H, int W, int C, int OH, int OW,
                                         cudaStream_t stream) {
  if (OH < 1 || OW < 1 || OH > H || OW > W || C < 1) return cudaErrorInvalidValue;
  if (C % 8 != 0) {
    const long long tot = (long long)B * OH * OW * C;
    const long long blk = (tot + 255) / 256;
    adaptive_avgpool_nhwc_scalar_kernel<<<(unsigned)(blk < 148 * 32 ? (blk < 1 ? 1 : blk) : 148 * 32), 256, 0, stream>>>(
        x, y, B, H, W, C, OH, OW);
    return cudaGetLastError();
  }
  const long long total = (long long)B * OH * OW * (C >> 3);
  const long long blocks = (total + 255) / 256;
  adaptive_avgpool_nhwc_kernel<<<(unsigned)(blocks < 148 * 32 ? (blocks < 1 ? 1 : blocks) : 148 * 32), 256, 0, stream>>>(
      x, y, B, H, W, C, OH, OW);
  return cudaGetLastError();
}

}  // namespace gdf
