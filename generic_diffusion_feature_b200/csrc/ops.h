// Internal op layer: every kernel of the extraction path behind a plain host function. The C ABI
// (gdf_api.cu) and the model executor (executor.cu) are both written against this header.
#pragma once
#include "gemm_sm100.cuh"
#include "host_util.h"

namespace gdf {

typedef __nv_bfloat16 bf16;

// A fully prepared GEMM launch: tensor maps + parameters. Built once at plan time, replayed per step.
struct GemmLaunch {
  GemmMaps maps;
  GemmParams p;
};

struct Epilogue {           // everything optional; zero-initialise then fill
  float alpha = 1.f;
  int n_out = 0;              // valid output columns; 0 = all (N, or N/2 for GEGLU)
  const float* bias = nullptr;
  const float* bias_m = nullptr;
  const float* row_batch_bias = nullptr;
  int rows_per_batch = 0;
  int act = kActNone;
  const float* col_scale = nullptr;
  const bf16* residual = nullptr; int ld_res = 0;
  float out_scale = 1.f;
  bf16* out = nullptr; int ld_out = 0; long long out_batch_stride = 0;
  int out_f16_from = 0;       // > 0: columns >= this go to `out` as fp16 (multiple of 32); < 0: all columns fp16
  bf16* out2 = nullptr; int ld_out2 = 0;
  float* out_f32 = nullptr; int ld_out_f32 = 0;
  __half* cap_pre = nullptr; int ld_cap_pre = 0;
  CaptureSeg cap[3] = {};
  int num_cap = 0;
  const float* ln_sums = nullptr;   // folded LayerNorm of the A rows (see GemmParams)
  const float* ln_u = nullptr;
  float ln_eps = 1e-5f;
  float* row_sums = nullptr;        // (sum, sum sq) of the final output rows for a LayerNorm folded into the consumer
  float* gn_sums = nullptr;         // fused GroupNorm statistics of the output (see GemmParams); groups of gn_cpg
  int gn_cpg = 0, gn_groups = 0;    //   channels (4 / 8 / 16), gn_rows_per_img rows (pixels) per image
  long long gn_rows_per_img = 0;
  bool in_f16 = false;              // operands are fp16 instead of bf16
  bool res_f16 = false;             // residual holds fp16 (captured feature map) instead of bf16: general epilogue
  // K-split of the last partial wave (GemmParams::sk_*): workspace of the caller (one launch at a time may use it: the
  // executor's op list runs on one stream). Null: whole tiles only. sk_cnt must be zero before the first launch.
  float* sk_ws = nullptr;
  long long sk_ws_floats = 0;
  unsigned int* sk_cnt = nullptr;
  int sk_cnt_len = 0;
  bool defer_capture_maps = false;  // capture pointers are placeholders: maps are built later (build_capture_maps)
};

int choose_block_n(int N, bool geglu, int num_m_tiles, int num_k_blocks = 0, bool k_split = false);
// K-split plan for a tail of `tail` tiles on `groups` resident CTA groups: pieces per tile (1 = do not split), k-blocks
// per piece, and the cost of the tail in k-block times (num_k_blocks when not split).
int plan_k_split(int num_k_blocks, int groups, int tail, int* kpp, float* tail_cost);
int gemm_resident_groups(int cta_group);   // CTA groups (CTAs or CTA pairs) one launch keeps resident

// C[M,N] = A[M,K] * W[N,K]^T, optionally batched (A: batch x M x K with a_batch_stride elements,
// W shared when w_batch_stride == 0).
int build_linear(GemmLaunch* g, const bf16* A, long long M, int K, int lda, const bf16* W, int N, int ldw,
                 const Epilogue& e, int batch = 1, long long a_batch_stride = 0, long long w_batch_stride = 0,
                 int block_n = 0);

// 3x3 convolution, NHWC bf16 contiguous input X[B, Hin, Win, Cin], packed weights Wp[Cout_padded][9*Cin]
// (k = (ky*3+kx)*Cin + c). stride 1: pad 1. stride 2: pad_lo 1 (UNet, symmetric pad 1) or 0 (VAE, pad (0,1,0,1)).
// Output rows are pixels of the output grid in (b, y, x) order; N = Cout (accumulator columns, multiple of 16).
int build_conv3x3(GemmLaunch* g, const bf16* X, int B, int Hin, int Win, int Cin, const bf16* Wp, int N, int stride,
                  int pad_lo, const Epilogue& e, int block_n = 0);

// (Re)builds the TMA-store tensor maps of the capture destinations (cap_pre, cap[0..2]) from the pointers currently
// in g->p; the executor calls it at run time once the arena base is known. No-op when g->p.tma_store == 0.
int build_capture_maps(GemmLaunch* g);

cudaError_t launch_gemm(const GemmMaps& maps, const GemmParams& p, cudaStream_t stream);
inline cudaError_t launch_gemm(const GemmLaunch& g, cudaStream_t s) { return launch_gemm(g.maps, g.p, s); }
int gemm_num_sms();

// ---- first VAE convolution fused from the fp32 NCHW image (conv_in_sm100.cu): no im2col operand in HBM
bool conv_in_fused_supported(int Cin, int N, int W);
int launch_conv_in_fused(const float* img, const bf16* w_packed, const float* bias, bf16* out, int B, int H, int W, int N,
                         float* gn_sums, int gn_cpg, int gn_groups, cudaStream_t stream);

// ---- normalisation (norm.cu)
// GroupNorm(+SiLU) over NHWC bf16 x[B, HW, C] -> y[B, HW, C]; stats in fp32, workspace >= gn_workspace_floats().
size_t gn_workspace_floats(int B, int G);
cudaError_t launch_groupnorm(const bf16* x, bf16* y, const float* gamma, const float* beta, int B, int HW, int C, int G,
                             float eps, bool silu, float* workspace, cudaStream_t stream);
// GroupNorm whose (sum, sum sq) per (image, group) were accumulated by the producing GEMM (Epilogue::gn_sums)
cudaError_t launch_groupnorm_from_sums(const bf16* x, bf16* y, const float* gamma, const float* beta, int B, int HW,
                                       int C, int G, float eps, bool silu, const float* sums, float* workspace,
                                       cudaStream_t stream);
// LayerNorm over rows of x[M, C] (ld = C) with optional affine and optional per-sample modulation
// y = LN(x) * (1 + scale[b]) + shift[b] (PixArt AdaLN-single), rows_per_batch rows per sample.
cudaError_t launch_layernorm(const bf16* x, bf16* y, const float* gamma, const float* beta, long long M, int C,
                             float eps, const float* mod_scale, const float* mod_shift, int rows_per_batch,
                             cudaStream_t stream);

// ---- attention (attention.cu): softmax(Q K^T * scale) V per (batch, head), head_dim 64.
// Q[B*Nq, ldq] / K,V[B*Nk, ldk/ldv] / O[B*Nq, ldo]; head h lives in columns [h*64, h*64+64).
// v_f16: V holds fp16 bit patterns (written by the projection GEMM with out_f16_from): enables the tcgen05 kernel,
// whose P.V product runs in fp16 (P = exp2 computed two-per-MUFU-op in half2).
cudaError_t launch_attention64(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo,
                               int B, int heads, int Nq, int Nk, float scale, int v_f16, cudaStream_t stream);
bool attention_uses_tcgen05(int Nk);
// head dims other than 64 (multiple of 8, <= 160): mma.sync kernel templated on the padded head dim
// key_bias (optional): fp32 [B, Nk] added to the scaled scores (PixArt's (1 - mask) * -10000 cross-attention bias)
cudaError_t launch_attention_generic(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O,
                                     int ldo, int B, int heads, int Nq, int Nk, int D, float scale,
                                     cudaStream_t stream, const float* key_bias = nullptr);
// Attention with the probability matrix as an output (attention.cu): P fp16 [B, heads, Nq, Nk] = softmax(scale Q K^T),
// O = P V (the reference's AttnStoreProcessor slow path); v_f16: V holds fp16 bit patterns. head dim % 8 == 0, <= 256.
cudaError_t launch_attention_probs(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, int v_f16,
                                   bf16* O, int ldo, __half* P, int B, int heads, int Nq, int Nk, int D, float scale,
                                   cudaStream_t stream, const float* key_bias = nullptr, __half* P2 = nullptr,
                                   int split = 0);
cudaError_t launch_head_mean(const __half* P, __half* out, int B, int heads, long long n, cudaStream_t stream);
// tensor-core form of the map path: fp32 scores -> fp16 probabilities; V of one image -> V^T fp16 [heads][D][Nk]
cudaError_t launch_softmax_rows_f32_f16(const float* S, __half* P, long long rows, int cols, cudaStream_t stream);
cudaError_t launch_transpose_v_f16(const bf16* V, int ldv, __half* VT, int heads, int Nk, int D, cudaStream_t stream);
// tcgen05 / TMEM flash attention (attention_sm100.cu); launch_attention64 dispatches to it for Nk >= 128.
int launch_attention64_tcgen05(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo,
                               int B, int heads, int Nq, int Nk, float scale, cudaStream_t stream);
// persistent tcgen05 / TMEM flash attention for head dims 40 / 64 / 72 / 80 / 128 and any key count (attention_tc.cu);
// v_f16: V (and P) are fp16 bit patterns, otherwise bf16. key_bias: optional fp32 [B, Nk] added to the scaled scores.
bool attention_tc_supports(int D);
void attention_tc_set_trace(void* buf, int cap);   // debug: device buffer of `cap` u64 that CTA 0 of the next launches fills
int launch_attention_tc(const bf16* Q, int ldq, const bf16* K, int ldk, const bf16* V, int ldv, bf16* O, int ldo, int B,
                        int heads, int Nq, int Nk, int D, float scale, int v_f16, const float* key_bias,
                        cudaStream_t stream);
// Row softmax in place over bf16 S[rows, cols] (ld), fp32 math (VAE single-head attention).
cudaError_t launch_softmax_rows(bf16* S, long long rows, int cols, int ld, cudaStream_t stream);

// ---- elementwise / data movement (eltwise.cu)
cudaError_t launch_upsample_nearest2x(const bf16* x, bf16* y, int B, int H, int W, int C, cudaStream_t stream);
// im2col for tiny Cin (3 or 4): X -> A[M, 64] bf16 with k = (ky*3+kx)*Cin + c, zero padded to 64.
//   src_nchw_f32: image (B, Cin, H, W) fp32;  src_nhwc_bf16: (B, H, W, Cin) bf16. Exactly one is non-null.
cudaError_t launch_im2col_small(const float* src_nchw_f32, const bf16* src_nhwc_bf16, bf16* A, int B, int H, int W,
                                int Cin, cudaStream_t stream);
// Posterior sample + scaling + q_sample + scale_model_input (reference: pipeline_pixart_sigma.py:644-673,
// diffusion_feature.py:406). moments: NHWC fp32 [B, HW, 8] (mean 0..3, logvar 4..7); eps_*: NCHW fp32 (B,4,h,w).
// latent_channels = LC (4; 16 for the Flux VAE): moments [B, HW, 2*LC]; shift_factor is subtracted before scaling (Flux).
cudaError_t launch_qsample(const float* moments, const float* eps_vae, const float* eps_q, float scaling_factor,
                           float shift_factor, float sqrt_ab, float sqrt_1m_ab, float input_scale, bf16* latent_nhwc,
                           __half* cap_unet_in, float* latents_nchw_f32, int B, int HW, int latent_channels,
                           cudaStream_t stream);
cudaError_t launch_cast_f32_to_bf16(const float* x, bf16* y, long long n, cudaStream_t stream);
cudaError_t launch_cast_bf16_to_f16(const bf16* x, __half* y, long long n, cudaStream_t stream);
// fp32 OIHW conv weight -> bf16 [O_pad][kh*kw*I] (k = (ky*kw+kx)*I + c), rows >= O zero.
cudaError_t launch_pack_conv_weight(const float* w_oihw, bf16* out, int O, int O_pad, int I, int kh, int kw,
                                    int k_pad, cudaStream_t stream);
cudaError_t launch_pack_conv_weight_f16(const float* w_oihw, __half* out, int O, int O_pad, int I, int kh, int kw,
                                        int k_pad, cudaStream_t stream);
// Sinusoidal timestep embedding (flip_sin_to_cos=True, freq_shift=0): out[n, dim] = [cos | sin].
cudaError_t launch_timestep_embedding(const float* t, float* out, int n, int dim, cudaStream_t stream);
// small fp32 GEMV-style linear for the conditioning MLPs: y[B, N] = act(x[B, K]) W[N, K]^T + b
struct GroupedLinearItem {   // one projection of a grouped GEMV (launch_grouped_small_linear)
  const float* W;            // [N, K] fp32
  const float* bias;         // [N] or null
  float* y;                  // [B, N]
  int N;
  int row0;                  // first row of this projection in the concatenated row space
};
cudaError_t launch_grouped_small_linear(const GroupedLinearItem* items_dev, int n_items, const float* x, int B, int K,
                                        int total_rows, int act_in_silu, cudaStream_t stream);
cudaError_t launch_small_linear(const float* x, const float* W, const float* b, float* y, int B, int K, int N,
                                int act_in_silu, int act_out_silu, cudaStream_t stream);

// PixArt DiT helpers (eltwise.cu)
cudaError_t launch_patchify(const bf16* x, bf16* A, int B, int L, int p, int Cin, int k_pad, cudaStream_t stream,
                            int chan_major = 0);
// Flux MMDiT helpers (eltwise.cu)
cudaError_t launch_qk_rmsnorm_rope(bf16* qkv, int ld, int rows, int heads, int hd, int k_off, const float* wq_a,
                                   const float* wk_a, const float* wq_b, const float* wk_b, int rows_a,
                                   const float* cos_t, const float* sin_t, float eps, cudaStream_t stream);
cudaError_t launch_copy_rows_bf16_f16(const bf16* src, int ld_src, __half* dst, int ld_dst, long long rows, int cols,
                                      cudaStream_t stream);
cudaError_t launch_sum3_f32(float* dst, const float* a, const float* b, const float* c, int n, cudaStream_t stream);
cudaError_t launch_replicate_rows_bf16(const float* src, bf16* dst, long long n, int B, cudaStream_t stream);
// out[j][b][c] = table[j][c] + t[b][j*C + c] (t_ld = J*C) or + t[b][c] (t_ld = C)
cudaError_t launch_adaln_mod(const float* table, const float* t, float* out, int B, int J, int C, int t_ld,
                             cudaStream_t stream);
cudaError_t launch_unpatchify(const float* x, float* out, int B, int g, int p, int oc, cudaStream_t stream);
cudaError_t launch_mask_to_bias(const float* mask, float* bias, int n, cudaStream_t stream);

// ---- feature_resize: adaptive average pooling of a captured fp16 NHWC map (stack.cu)
cudaError_t launch_adaptive_avgpool_nhwc(const __half* x, __half* y, int B, int H, int W, int C, int OH, int OW,
                                         cudaStream_t stream);

// ---- feature stack + correspondence (stack.cu)
struct ResizeSrc {
  const __half* ptr;  // fp16 NHWC map [B, h*w, C]
  int h, w, C, c_off; // c_off: channel offset inside the stack
};
// Bilinear (align_corners=False) resize of every map to (OH, OW) + channel concat. out_nhwc: [B, OH*OW, Ctot]
// (fp16); out_nchw: [B, Ctot, OH, OW] (reference layout). sumsq (optional): [B, OH*OW] fp32 accumulated
// squared L2 norm per pixel of the NHWC stack.
// GDF_DETERMINISTIC=1: no floating-point atomics anywhere on the path (GroupNorm statistics by their own kernel with a
// fixed-order reduction instead of the producing GEMM's epilogue): two runs give bit-identical features.
bool deterministic_mode();
cudaError_t launch_resize_concat(const ResizeSrc* srcs_host, int n_src, int B, int OH, int OW, int Ctot,
                                 __half* out_nhwc, __half* out_nchw, float* sumsq, cudaStream_t stream);

// ---- correspondence (stack.cu): cosine-similarity arg-max of n query points against a target stack, evaluated in
// the exact low-resolution form of correspondence_utils.find_nn_source_correspondences (:113-138):
//   sims_up(i, p) = bilinear_p( q_i . F2 ) / | bilinear_p(F2) |   (bilinear interpolation is linear)
// stack1 / stack2: fp16 NHWC [hw*hw, C]; query_yx: int32 (n, 2) positions on the load x load grid;
// idx_out: int64 (n) flat index into the load x load grid. workspace floats: corr_workspace_floats().
size_t corr_workspace_floats(int n, int hw, int C);
int launch_correspond(const __half* stack1, const __half* stack2, int C, int hw, int load, const int* query_yx, int n,
                      long long* idx_out, float* workspace, cudaStream_t stream);

}  // namespace gdf
