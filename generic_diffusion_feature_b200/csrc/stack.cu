// Feature-stack construction: bilinear resize (align_corners=False, PyTorch semantics) of captured fp16 maps
// + channel concat, in NHWC (pixel-major, feeds the similarity GEMM) and/or NCHW (the reference's layout).
// Reference: F.interpolate(f, (128,128), mode='bilinear') + torch.cat in
// correspondence/correspondence/aggregation_network.py:62-66. HBM-bound; NCHW output is transposed through
// shared memory so both the reads (channel-contiguous) and the writes (pixel-contiguous) are coalesced.
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

// NHWC -> NCHW through shared memory: block = 32 output pixels (one row segment) x 64 channels.
__global__ void __launch_bounds__(256)
resize_nchw_kernel(const __half* __restrict__ src, int h, int w, int C, int c_off, int B, int OH, int OW, int Ctot,
                   __half* __restrict__ out) {
  __shared__ __half tile[64][32 + 2];
  const int xblocks = (OW + 31) / 32;
  const int xb = blockIdx.x % xblocks;
  const int oy = (blockIdx.x / xblocks) % OH;
  const int b = blockIdx.x / (xblocks * OH);
  const int cb = blockIdx.y * 64;
  const float sy = (float)h / (float)OH, sx = (float)w / (float)OW;
  const int C8 = C / 8;
  {  // 32 pixels x 8 channel-vectors = 256 work items
    const int px = threadIdx.x >> 3, cv = threadIdx.x & 7;
    const int ox = xb * 32 + px;
    const int c8 = cb / 8 + cv;
    if (ox < OW && c8 < C8) {
      const BilinearTap ty = bilinear_tap(oy, h, sy), tx = bilinear_tap(ox, w, sx);
      const uint4* s = reinterpret_cast<const uint4*>(src + ((long long)b * h * w) * C) + c8;
      const uint4 v00 = __ldg(s + ((long long)ty.i0 * w + tx.i0) * C8);
      const uint4 v01 = __ldg(s + ((long long)ty.i0 * w + tx.i1) * C8);
      const uint4 v10 = __ldg(s + ((long long)ty.i1 * w + tx.i0) * C8);
      const uint4 v11 = __ldg(s + ((long long)ty.i1 * w + tx.i1) * C8);
      float o[8];
      blend8(v00, v01, v10, v11, ty.w0, ty.w1, tx.w0, tx.w1, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) tile[cv * 8 + j][px] = __float2half_rn(o[j]);
    }
  }
  __syncthreads();
  {  // 64 channels x 32 pixels, pixel fastest
    for (int i = threadIdx.x; i < 64 * 32; i += 256) {
      const int ch = i >> 5, px = i & 31;
      const int ox = xb * 32 + px;
      if (ox < OW && cb + ch < C)
        out[(((long long)b * Ctot + c_off + cb + ch) * OH + oy) * OW + ox] = tile[ch][px];
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
  if (Ctot % 8 != 0) return cudaErrorInvalidValue;
  for (int i = 0; i < n_src; ++i) {
    const ResizeSrc& s = srcs_host[i];
    if (s.C % 8 != 0 || s.c_off % 8 != 0) return cudaErrorInvalidValue;
    if (out_nhwc) {
      const long long total = (long long)B * OH * OW * (s.C / 8);
      const long long blocks = (total + 255) / 256;
      resize_nhwc_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, stream>>>(
          s.ptr, s.h, s.w, s.C, s.c_off, B, OH, OW, Ctot, out_nhwc);
    }
    if (out_nchw) {
      dim3 grid((unsigned)(((OW + 31) / 32) * OH * B), (unsigned)((s.C + 63) / 64));
      resize_nchw_kernel<<<grid, 256, 0, stream>>>(s.ptr, s.h, s.w, s.C, s.c_off, B, OH, OW, Ctot, out_nchw);
    }
  }
  if (sumsq && out_nhwc) {
    const long long rows = (long long)B * OH * OW;
    rownorm_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(out_nhwc, rows, Ctot, sumsq);
  }
  return cudaGetLastError();
}

}  // namespace gdf
