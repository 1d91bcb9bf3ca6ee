// Bandwidth-bound helpers around the GEMMs: q_sample, nearest upsample, tiny-Cin im2col, dtype casts,
// weight packing and the conditioning-embedding math (sinusoids + small fp32 linears).
#include <type_traits>
#include "ops.h"

namespace gdf {

// ---------------------------------------------------------------- nearest 2x upsample (NHWC bf16)
// Reference: F.interpolate(scale_factor=2.0, mode="nearest") in feature/diffusers/models/upsampling.py:176-179.
__global__ void upsample_nearest2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                          int C8) {
  const long long total = (long long)B * (2 * H) * (2 * W) * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int ox = (int)(r % (2 * W));
    r /= (2 * W);
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    y[i] = __ldg(x + (((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * C8 + c);
  }
}
cudaError_t launch_upsample_nearest2x(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t stream) {
  if (C % 8 != 0) return cudaErrorInvalidValue;
  const long long total = (long long)B * 4 * H * W * (C / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  upsample_nearest2x_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                        reinterpret_cast<uint4*>(y), B, H, W, C / 8);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- im2col for Cin in {3, 4}
// One thread per output pixel writes one 128-byte row A[m, 0..63] (k = (ky*3+kx)*Cin + c, zero padded).
__global__ void im2col_small_kernel(const float* __restrict__ src_f32, const bf16* __restrict__ src_bf16,
                                    bf16* __restrict__ A, int B, int H, int W, int Cin) {
  const long long total = (long long)B * H * W;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total;
       m += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(m % W);
    long long r = m / W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float v[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) v[k] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = y + ky - 1, ix = x + kx - 1;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        for (int c = 0; c < Cin; ++c) {
          float t;
          if (src_f32) t = __ldg(src_f32 + (((long long)b * Cin + c) * H + iy) * W + ix);
          else t = __bfloat162float(src_bf16[(((long long)b * H + iy) * W + ix) * Cin + c]);
          v[(ky * 3 + kx) * Cin + c] = t;
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(A + m * 64);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 o;
      o.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
      o.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
      o.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
      o.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
      dst[q] = o;
    }
  }
}
cudaError_t launch_im2col_small(const float* src_nchw_f32, const bf16* src_nhwc_bf16, bf16* A, int B, int H, int W,
                                int Cin, cudaStream_t stream) {
  if (Cin * 9 > 64 || (!src_nchw_f32 == !src_nhwc_bf16)) return cudaErrorInvalidValue;
  const long long total = (long long)B * H * W;
  const int blocks = (int)((total + 127) / 128 < 148 * 32 ? (total + 127) / 128 : 148 * 32);
  im2col_small_kernel<<<blocks, 128, 0, stream>>>(src_nchw_f32, src_nhwc_bf16, A, B, H, W, Cin);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- posterior sample + q_sample
// z = (mean + exp(0.5*clamp(logvar,-30,20)) * eps_vae - shift_factor) * scaling_factor   (DiagonalGaussianDistribution
// .sample; shift_factor != 0 only for the Flux VAE, pipeline_flux_img2img.py _encode_vae_image)
// x_t = sqrt_ab * z + sqrt_1m_ab * eps_q                                      (scheduler.add_noise)
// model input = x_t * input_scale                                            (scheduler.scale_model_input)
__global__ void qsample_kernel(const float* __restrict__ moments, const float* __restrict__ eps_vae,
                               const float* __restrict__ eps_q, float sf, float shift, float sqrt_ab, float sqrt_1m_ab,
                               float input_scale, bf16* __restrict__ latent, __half* __restrict__ cap,
                               float* __restrict__ latents_nchw, int B, int HW, int LC) {
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int p = (int)(i % HW);
    const float* mo = moments + i * 2 * LC;
    for (int c = 0; c < LC; ++c) {
      const float mean = mo[c];
      const float logvar = fminf(fmaxf(mo[LC + c], -30.f), 20.f);
      const long long ni = ((long long)b * LC + c) * HW + p;
      const float z = (mean + expf(0.5f * logvar) * eps_vae[ni] - shift) * sf;
      const float xt = sqrt_ab * z + sqrt_1m_ab * eps_q[ni];
      const float xin = xt * input_scale;
      if (latents_nchw) latents_nchw[ni] = xt;
      latent[i * LC + c] = __float2bfloat16_rn(xin);
      if (cap) cap[i * LC + c] = __float2half_rn(xin);
    }
  }
}
cudaError_t launch_qsample(const float* moments, const float* eps_vae, const float* eps_q, float scaling_factor,
                           float shift_factor, float sqrt_ab, float sqrt_1m_ab, float input_scale, bf16* latent_nhwc,
                           __half* cap_unet_in, float* latents_nchw_f32, int B, int HW, int latent_channels,
                           cudaStream_t stream) {
  const long long total = (long long)B * HW;
  const int blocks = (int)((total + 255) / 256);
  qsample_kernel<<<blocks, 256, 0, stream>>>(moments, eps_vae, eps_q, scaling_factor, shift_factor, sqrt_ab,
                                             sqrt_1m_ab, input_scale, latent_nhwc, cap_unet_in, latents_nchw_f32, B,
                                             HW, latent_channels);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- casts
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void cast_bf16_f16_kernel(const bf16* __restrict__ x, __half* __restrict__ y, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2half_rn(__bfloat162float(x[i]));
}
static int grid_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 148 * 32 ? (b < 1 ? 1 : b) : 148 * 32);
}
cudaError_t launch_cast_f32_to_bf16(const float* x, bf16* y, long long n, cudaStream_t stream) {
  cast_f32_bf16_kernel<<<grid_for(n), 256, 0, stream>>>(x, y, n);
  return cudaGetLastError();
}
cudaError_t launch_cast_bf16_to_f16(const bf16* x, __half* y, long long n, cudaStream_t stream) {
  cast_bf16_f16_kernel<<<grid_for(n), 256, 0, stream>>>(x, y, n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- conv weight packing
// w[O][I][kh][kw] fp32 -> out[O_pad][k_pad] bf16, k = (ky*kw+kx)*I + c, zero padding rows/cols.
template <typename T>
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int O, int O_pad, int I,
                                        int kh, int kw, int k_pad) {
  const long long total = (long long)O_pad * k_pad;
  const int Kreal = kh * kw * I;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % k_pad);
    const int o = (int)(i / k_pad);
    float v = 0.f;
    if (o < O && k < Kreal) {
      const int c = k % I;
      const int tap = k / I;
      const int ky = tap / kw, kx = tap % kw;
      v = w[(((long long)o * I + c) * kh + ky) * kw + kx];
    }
    if constexpr (sizeof(T) == 2 && std::is_same<T, __half>::value) out[i] = __float2half_rn(v);
    else out[i] = __float2bfloat16_rn(v);
  }
}
cudaError_t launch_pack_conv_weight(const float* w_oihw, bf16* out, int O, int O_pad, int I, int kh, int kw, int k_pad,
                                    cudaStream_t stream) {
  pack_conv_weight_kernel<bf16><<<grid_for((long long)O_pad * k_pad), 256, 0, stream>>>(w_oihw, out, O, O_pad, I, kh,
                                                                                      kw, k_pad);
  return cudaGetLastError();
}
// same packing in fp16: weights of a convolution over an fp16 tensor (feature stacks; kind::f16 needs A and B of one type)
cudaError_t launch_pack_conv_weight_f16(const float* w_oihw, __half* out, int O, int O_pad, int I, int kh, int kw,
                                        int k_pad, cudaStream_t stream) {
  pack_conv_weight_kernel<__half><<<grid_for((long long)O_pad * k_pad), 256, 0, stream>>>(w_oihw, out, O, O_pad, I, kh,
                                                                                        kw, k_pad);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- conditioning embeddings (fp32, tiny)
// Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): emb = t * exp(-ln(10000) * i / half), out = [cos|sin]
__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int n, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  const float freq = expf(-9.210340371976184f * (float)j / (float)half);
  const float a = t[r] * freq;
  out[(long long)r * dim + j] = cosf(a);
  out[(long long)r * dim + half + j] = sinf(a);
}
cudaError_t launch_timestep_embedding(const float* t, float* out, int n, int dim, cudaStream_t stream) {
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, stream>>>(t, out, n, dim);
  return cudaGetLastError();
}

// y[b, n] = (silu_in ? silu(x) : x)[b, :] . W[n, :] + bias[n]; one warp per output COLUMN n, all B rows at once (the
// weight row is streamed once with 128-bit loads and reused for every batch row; B <= 8 per pass). These GEMVs are
// pure weight streaming: 57 Flux modulation projections read 12.9 GB of fp32 weights per forward.
template <int kB>
__global__ void small_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ y, int B0, int B, int K, int N,
                                    int act_in_silu, int act_out_silu) {
  const int lane = threadIdx.x & 31;
  const long long n = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float* wr = W + n * K;
  float acc[kB];
#pragma unroll
  for (int b = 0; b < kB; ++b) acc[b] = 0.f;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        if (B0 + b < B) {
          float4 xv = *reinterpret_cast<const float4*>(x + (long long)(B0 + b) * K + k);
          if (act_in_silu) {
            xv.x = xv.x / (1.f + expf(-xv.x)); xv.y = xv.y / (1.f + expf(-xv.y));
            xv.z = xv.z / (1.f + expf(-xv.z)); xv.w = xv.w / (1.f + expf(-xv.w));
          }
          acc[b] += xv.x * w4.x + xv.y * w4.y + xv.z * w4.z + xv.w * w4.w;
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int b = 0; b < kB; ++b) {
        if (B0 + b < B) {
          float xv = x[(long long)(B0 + b) * K + k];
          if (act_in_silu) xv = xv / (1.f + expf(-xv));
          acc[b] += xv * wv;
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kB; ++b) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], s);
  }
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kB; ++b) {
      if (B0 + b < B) {
        float r = acc[b] + (bias ? bias[n] : 0.f);
        if (act_out_silu) r = r / (1.f + expf(-r));
        y[(long long)(B0 + b) * N + n] = r;
      }
    }
  }
}
// one warp per output ELEMENT (b, n): most parallelism, weights re-read per batch row (fine while they sit in L2)
__global__ void small_linear_elem_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                         const float* __restrict__ bias, float* __restrict__ y, int B, int K, int N,
                                         int act_in_silu, int act_out_silu) {
  const int lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= (long long)B * N) return;
  const int b = (int)(o / N), n = (int)(o % N);
  const float* xr = x + (long long)b * K;
  const float* wr = W + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = xr[k];
    if (act_in_silu) xv = xv / (1.f + expf(-xv));
    acc += xv * __ldg(wr + k);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    float r = acc + (bias ? bias[n] : 0.f);
    if (act_out_silu) r = r / (1.f + expf(-r));
    y[o] = r;
  }
}
// Grouped GEMV: y_i[b, n] = bias_i[n] + sum_k act(x[b, k]) W_i[n, k] for a table of (W_i, bias_i, y_i, N_i) that all read
// the same x (the time-embedding projections of every resnet of a UNet, resnet.py:354-355: Linear(SiLU(temb))). One
// launch at the head of the op list instead of one cold fp32 GEMV per resnet inside the dependency chain: one warp per
// output row of the concatenated problem, the B rows of x staged in shared memory, every weight row read once.
__global__ void __launch_bounds__(256)
grouped_small_linear_kernel(const GroupedLinearItem* __restrict__ items, int n_items, const float* __restrict__ x, int B,
                            int K, int total_rows, int act_in_silu) {
  extern __shared__ float gsl_x[];   // [B][K]
  for (int i = threadIdx.x; i < B * K; i += blockDim.x) {
    float v = x[i];
    if (act_in_silu) v = v / (1.f + expf(-v));
    gsl_x[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= total_rows) return;
  int it = 0;
  while (it + 1 < n_items && items[it + 1].row0 <= r) ++it;   // <= 64 groups: a linear scan of a table in L1
  const GroupedLinearItem g = items[it];
  const int n = r - g.row0;
  const float* wr = g.W + (long long)n * K;
  for (int b0 = 0; b0 < B; b0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float w = __ldg(wr + k);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (b0 + j < B) acc[j] = fmaf(w, gsl_x[(b0 + j) * K + k], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = acc[j];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
      if (lane == 0 && b0 + j < B) g.y[(long long)(b0 + j) * g.N + n] = a + (g.bias ? g.bias[n] : 0.f);
    }
  }
}
cudaError_t launch_grouped_small_linear(const GroupedLinearItem* items_dev, int n_items, const float* x, int B, int K,
                                        int total_rows, int act_in_silu, cudaStream_t stream) {
  const size_t smem = (size_t)B * K * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;   // callers fall back to one launch per projection
  grouped_small_linear_kernel<<<(unsigned)((total_rows + 7) / 8), 256, smem, stream>>>(items_dev, n_items, x, B, K,
                                                                                      total_rows, act_in_silu);
  return cudaGetLastError();
}

cudaError_t launch_small_linear(const float* x, const float* W, const float* b, float* y, int B, int K, int N,
                                int act_in_silu, int act_out_silu, cudaStream_t stream) {
  if ((long long)N * K * 4 <= (32ll << 20)) {   // L2-resident weights (time / text embedding MLPs): favour parallelism
    const long long outs = (long long)B * N;
    small_linear_elem_kernel<<<(unsigned)((outs + 7) / 8), 256, 0, stream>>>(x, W, b, y, B, K, N, act_in_silu,
                                                                            act_out_silu);
    return cudaGetLastError();
  }
  const unsigned blocks = (unsigned)((N + 7) / 8);
  for (int b0 = 0; b0 < B; b0 += 8) {   // weight streaming (Flux modulation projections): 8 batch rows per pass
    const int nb = B - b0;
    if (nb == 1)
      small_linear_kernel<1><<<blocks, 256, 0, stream>>>(x, W, b, y, b0, B, K, N, act_in_silu, act_out_silu);
    else if (nb == 2)
      small_linear_kernel<2><<<blocks, 256, 0, stream>>>(x, W, b, y, b0, B, K, N, act_in_silu, act_out_silu);
    else if (nb <= 4)
      small_linear_kernel<4><<<blocks, 256, 0, stream>>>(x, W, b, y, b0, B, K, N, act_in_silu, act_out_silu);
    else
      small_linear_kernel<8><<<blocks, 256, 0, stream>>>(x, W, b, y, b0, B, K, N, act_in_silu, act_out_silu);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- PixArt DiT helpers
// PatchEmbed im2col (Conv2d k = s = p, [diffusers embeddings.PatchEmbed]; reference use: transformer_2d.py:541-569):
// latent NHWC bf16 [B, L, L, Cin] -> A[B*(L/p)^2, k_pad] with k = (py*p + px)*Cin + c, zero padded.
// chan_major = 1: k = c*p*p + py*p + px (FluxImg2ImgPipeline._pack_latents: view(B,C,h/2,2,w/2,2).permute(0,2,4,1,3,5))
__global__ void patchify_kernel(const bf16* __restrict__ x, bf16* __restrict__ A, int B, int L, int p, int Cin,
                                int k_pad, int chan_major) {
  const int g = L / p;
  const long long total = (long long)B * g * g * k_pad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % k_pad);
    long long m = i / k_pad;
    const int gx = (int)(m % g);
    m /= g;
    const int gy = (int)(m % g);
    const int b = (int)(m / g);
    bf16 v = __float2bfloat16_rn(0.f);
    if (k < p * p * Cin) {
      const int c = chan_major ? k / (p * p) : k % Cin, pp = chan_major ? k % (p * p) : k / Cin;
      const int py = pp / p, px = pp % p;
      v = x[(((long long)b * L + gy * p + py) * L + gx * p + px) * Cin + c];
    }
    A[i] = v;
  }
}
cudaError_t launch_patchify(const bf16* x, bf16* A, int B, int L, int p, int Cin, int k_pad, cudaStream_t stream,
                            int chan_major) {
  if (L % p != 0 || p * p * Cin > k_pad) return cudaErrorInvalidValue;
  const long long total = (long long)B * (L / p) * (L / p) * k_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  patchify_kernel<<<blocks, 256, 0, stream>>>(x, A, B, L, p, Cin, k_pad, chan_major);
  return cudaGetLastError();
}
// fp32 [N, C] table -> bf16 [B, N, C] (position embedding replicated over the batch: residual operand of the
// patch-embedding GEMM)
__global__ void replicate_rows_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n, int B) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * B;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(src[i % n]);
}
cudaError_t launch_replicate_rows_bf16(const float* src, bf16* dst, long long n, int B, cudaStream_t stream) {
  replicate_rows_bf16_kernel<<<148 * 8, 256, 0, stream>>>(src, dst, n, B);
  return cudaGetLastError();
}
// AdaLN-single modulation vectors (attention.py:498-500): out[j][b][c] = table[j][c] + t[b][j*C + c], j < J
__global__ void adaln_mod_kernel(const float* __restrict__ table, const float* __restrict__ t, float* __restrict__ out,
                                 int B, int J, int C, int t_ld) {
  const int total = J * B * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % C, b = (i / C) % B, j = i / (C * B);
    out[i] = table[j * C + c] + t[(long long)b * t_ld + (t_ld == C ? 0 : j * C) + c];
  }
}
cudaError_t launch_adaln_mod(const float* table, const float* t, float* out, int B, int J, int C, int t_ld,
                             cudaStream_t stream) {
  const int total = J * B * C;
  adaln_mod_kernel<<<(total + 255) / 256, 256, 0, stream>>>(table, t, out, B, J, C, t_ld);
  return cudaGetLastError();
}
// unpatchify ([pixart_transformer_2d.py forward tail]: reshape (n,h,w,p,q,c) -> einsum nhwpqc->nchpwq):
// x fp32 [B*g*g, p*p*oc] (column = (py*p + px)*oc + c) -> out fp32 NCHW (B, oc, g*p, g*p)
__global__ void unpatchify_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int g, int p, int oc) {
  const int S = g * p;
  const long long total = (long long)B * oc * S * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % S);
    long long r = i / S;
    const int Y = (int)(r % S);
    r /= S;
    const int c = (int)(r % oc), b = (int)(r / oc);
    const int gy = Y / p, py = Y % p, gx = X / p, px = X % p;
    out[i] = x[(((long long)b * g + gy) * g + gx) * (p * p * oc) + (py * p + px) * oc + c];
  }
}
cudaError_t launch_unpatchify(const float* x, float* out, int B, int g, int p, int oc, cudaStream_t stream) {
  unpatchify_kernel<<<148 * 4, 256, 0, stream>>>(x, out, B, g, p, oc);
  return cudaGetLastError();
}
// encoder_attention_mask (1 = keep) -> additive key bias (1 - m) * -10000 ([pixart_transformer_2d.py forward head])
__global__ void mask_to_bias_kernel(const float* __restrict__ m, float* __restrict__ bias, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bias[i] = (1.f - m[i]) * -10000.f;
}
cudaError_t launch_mask_to_bias(const float* mask, float* bias, int n, cudaStream_t stream) {
  mask_to_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(mask, bias, n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- Flux MMDiT helpers
// RMS qk-norm + rotary embedding, in place on the fused QKV projection (FluxAttnProcessor2_0,
// attention_processor.py:2298-2301, 2319-2322 norm_q / norm_k / norm_added_q / norm_added_k and :2330-2334
// apply_rotary_emb; [diffusers normalization.RMSNorm, embeddings.apply_rotary_emb use_real_unbind_dim=-1, un-vendored]):
//   y = x * rsqrt(mean(x^2) + eps) * w            over the head_dim of one (token, head)
//   o[2i] = y[2i] cos[2i] - y[2i+1] sin[2i] ;  o[2i+1] = y[2i+1] cos[2i+1] + y[2i] sin[2i+1]
// qkv: bf16 [rows, ld], q in columns [0, heads*hd), k in [k_off, k_off + heads*hd). Rows < rows_a use (wq_a, wk_a)
// (the text stream's norm_added_* of a double block), the others (wq_b, wk_b). cos / sin: fp32 [rows, hd].
// One warp per (row, head, q|k); lanes own (even, odd) pairs, so the rotation needs no shuffles.
__global__ void qk_rmsnorm_rope_kernel(bf16* __restrict__ qkv, int ld, int rows, int heads, int hd, int k_off,
                                       const float* __restrict__ wq_a, const float* __restrict__ wk_a,
                                       const float* __restrict__ wq_b, const float* __restrict__ wk_b, int rows_a,
                                       const float* __restrict__ cos_t, const float* __restrict__ sin_t, float eps) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)rows * heads * 2) return;
  const int which = (int)(wid & 1);
  const int head = (int)((wid >> 1) % heads);
  const int row = (int)((wid >> 1) / heads);
  bf16* x = qkv + (long long)row * ld + (which ? k_off : 0) + head * hd;
  const float* w = row < rows_a ? (which ? wk_a : wq_a) : (which ? wk_b : wq_b);
  const int pairs = hd >> 1;
  float2 v[4];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = lane + 32 * j;
    v[j] = make_float2(0.f, 0.f);
    if (p < pairs) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(x + 2 * p);
      v[j] = __bfloat1622float2(t);
      ss = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, ss));
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  const float r = rsqrtf(ss / (float)hd + eps);
  const float* cr = cos_t ? cos_t + (long long)row * hd : nullptr;
  const float* sr = sin_t ? sin_t + (long long)row * hd : nullptr;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = lane + 32 * j;
    if (p < pairs) {
      const float y0 = v[j].x * r * (w ? w[2 * p] : 1.f), y1 = v[j].y * r * (w ? w[2 * p + 1] : 1.f);
      float o0 = y0, o1 = y1;
      if (cr) {
        o0 = y0 * cr[2 * p] - y1 * sr[2 * p];
        o1 = y1 * cr[2 * p + 1] + y0 * sr[2 * p + 1];
      }
      *reinterpret_cast<__nv_bfloat162*>(x + 2 * p) = __floats2bfloat162_rn(o0, o1);
    }
  }
}
cudaError_t launch_qk_rmsnorm_rope(bf16* qkv, int ld, int rows, int heads, int hd, int k_off, const float* wq_a,
                                   const float* wk_a, const float* wq_b, const float* wk_b, int rows_a,
                                   const float* cos_t, const float* sin_t, float eps, cudaStream_t stream) {
  if (hd % 2 != 0 || hd > 256 || ld % 2 != 0 || k_off % 2 != 0) return cudaErrorInvalidValue;
  const long long warps = (long long)rows * heads * 2;
  const int wpb = 8;
  qk_rmsnorm_rope_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      qkv, ld, rows, heads, hd, k_off, wq_a, wk_a, wq_b, wk_b, rows_a, cos_t, sin_t, eps);
  return cudaGetLastError();
}

// bf16 [rows, cols] block with row pitch ld_src -> fp16 with pitch ld_dst (captures of tensors that no GEMM epilogue
// produces: Flux `norm-out` / `out` = the modulated LayerNorm output, transformer_flux.py:200-211, and the single
// blocks' `attn-out` = the attention output's image rows, attention_processor.py:2358-2360). cols, pitches % 8 == 0.
__global__ void copy_rows_bf16_f16_kernel(const bf16* __restrict__ src, int ld_src, __half* __restrict__ dst, int ld_dst,
                                          long long rows, int cols) {
  const int cv = cols >> 3;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cv;
    const int c = (int)(i % cv) << 3;
    const uint4 u = *reinterpret_cast<const uint4*>(src + r * ld_src + c);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint4 o;
    __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) h2[q] = __float22half2_rn(__bfloat1622float2(b2[q]));
    *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = o;
  }
}
cudaError_t launch_copy_rows_bf16_f16(const bf16* src, int ld_src, __half* dst, int ld_dst, long long rows, int cols,
                                      cudaStream_t stream) {
  if (cols % 8 != 0 || ld_src % 8 != 0 || ld_dst % 8 != 0) return cudaErrorInvalidValue;
  copy_rows_bf16_f16_kernel<<<grid_for(rows * (cols >> 3)), 256, 0, stream>>>(src, ld_src, dst, ld_dst, rows, cols);
  return cudaGetLastError();
}

// dst = a + b (+ c): the three conditioning embeddings of CombinedTimestepGuidanceTextProjEmbeddings
__global__ void sum3_f32_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b,
                                const float* __restrict__ c, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + b[i] + (c ? c[i] : 0.f);
}
cudaError_t launch_sum3_f32(float* dst, const float* a, const float* b, const float* c, int n, cudaStream_t stream) {
  sum3_f32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dst, a, b, c, n);
  return cudaGetLastError();
}

}  // namespace gdf
