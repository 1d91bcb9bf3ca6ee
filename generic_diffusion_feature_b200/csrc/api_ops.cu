// C ABI, op level (include/gdf.h): thin extern "C" wrappers over ops.h.
#include "../../include/gdf.h"
#include <vector>
#include "ops.h"

using namespace gdf;

static Epilogue to_epilogue(const gdf_epilogue* ep) {
  Epilogue e;
  if (!ep) return e;
  e.alpha = ep->alpha == 0.f ? 1.f : ep->alpha;
  e.n_out = ep->n_out;
  e.bias = static_cast<const float*>(ep->bias_dev);
  e.bias_m = static_cast<const float*>(ep->bias_m_dev);
  e.row_batch_bias = static_cast<const float*>(ep->row_batch_bias_dev);
  e.rows_per_batch = ep->rows_per_batch;
  e.act = ep->act;
  e.col_scale = static_cast<const float*>(ep->col_scale_dev);
  e.residual = static_cast<const bf16*>(ep->residual_dev);
  e.ld_res = ep->ld_res;
  e.out_scale = ep->out_scale == 0.f ? 1.f : ep->out_scale;
  e.out = static_cast<bf16*>(ep->out_dev);
  e.ld_out = ep->ld_out;
  e.out_batch_stride = ep->out_batch_stride;
  e.out_f16_from = ep->out_f16_from;
  e.out2 = static_cast<bf16*>(ep->out2_dev);
  e.ld_out2 = ep->ld_out2;
  e.out_f32 = static_cast<float*>(ep->out_f32_dev);
  e.ld_out_f32 = ep->ld_out_f32;
  e.cap_pre = static_cast<__half*>(ep->cap_pre_dev);
  e.ld_cap_pre = ep->ld_cap_pre;
  e.num_cap = ep->num_cap;
  for (int i = 0; i < 3; ++i) {
    e.cap[i].ptr = static_cast<__half*>(ep->cap[i].ptr_dev);
    e.cap[i].col_begin = ep->cap[i].col_begin;
    e.cap[i].col_end = ep->cap[i].col_end;
    e.cap[i].ld = ep->cap[i].ld;
  }
  e.in_f16 = ep->in_f16 != 0;
  e.res_f16 = ep->res_f16 != 0;
  e.sk_ws = static_cast<float*>(ep->k_split_ws_dev);
  e.sk_ws_floats = ep->k_split_ws_floats;
  e.sk_cnt = static_cast<unsigned int*>(ep->k_split_cnt_dev);
  e.sk_cnt_len = ep->k_split_cnt_len;
  e.ln_sums = static_cast<const float*>(ep->ln_sums_dev);
  e.ln_u = static_cast<const float*>(ep->ln_u_dev);
  e.ln_eps = ep->ln_eps;
  e.row_sums = static_cast<float*>(ep->row_sums_dev);
  e.gn_sums = static_cast<float*>(ep->gn_sums_dev);
  e.gn_cpg = ep->gn_cpg;
  e.gn_groups = ep->gn_groups;
  e.gn_rows_per_img = ep->gn_rows_per_img;
  return e;
}

#define GDF_LAUNCH(expr)                                                                              \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return fail(GDF_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(_e));     \
    return GDF_OK;                                                                                    \
  } while (0)

extern "C" {

const char* gdf_last_error(void) { return last_error().c_str(); }
int gdf_abi_version(void) { return GDF_ABI_VERSION; }

int gdf_op_linear(const void* a_dev, int64_t M, int K, int lda, const void* w_dev, int N, int ldw,
                  const gdf_epilogue* ep, int batch, int64_t a_batch_stride, int64_t w_batch_stride, int block_n,
                  void* stream) {
  GemmLaunch g;
  GDF_TRY(build_linear(&g, static_cast<const bf16*>(a_dev), M, K, lda, static_cast<const bf16*>(w_dev), N, ldw,
                       to_epilogue(ep), batch < 1 ? 1 : batch, a_batch_stride, w_batch_stride, block_n));
  GDF_LAUNCH(launch_gemm(g, static_cast<cudaStream_t>(stream)));
}

int gdf_op_conv3x3(const void* x_dev, int B, int Hin, int Win, int Cin, const void* w_packed_dev, int N, int stride,
                   int pad_lo, const gdf_epilogue* ep, int block_n, void* stream) {
  GemmLaunch g;
  GDF_TRY(build_conv3x3(&g, static_cast<const bf16*>(x_dev), B, Hin, Win, Cin, static_cast<const bf16*>(w_packed_dev),
                        N, stride, pad_lo, to_epilogue(ep), block_n));
  GDF_LAUNCH(launch_gemm(g, static_cast<cudaStream_t>(stream)));
}

int gdf_op_conv_in(const void* img, const void* w_packed, const void* bias, void* out, int B, int H, int W, int N,
                   void* gn_sums, int gn_cpg, int gn_groups, void* stream) {
  return launch_conv_in_fused(static_cast<const float*>(img), static_cast<const bf16*>(w_packed),
                              static_cast<const float*>(bias), static_cast<bf16*>(out), B, H, W, N,
                              static_cast<float*>(gn_sums), gn_cpg, gn_groups, static_cast<cudaStream_t>(stream));
}

int gdf_op_pack_conv_weight_f16(const void* w, void* out, int O, int O_pad, int I, int kh, int kw, int k_pad,
                                void* stream) {
  GDF_LAUNCH(launch_pack_conv_weight_f16(static_cast<const float*>(w), static_cast<__half*>(out), O, O_pad, I, kh, kw,
                                         k_pad, static_cast<cudaStream_t>(stream)));
}

int gdf_op_pack_conv_weight(const void* w, void* out, int O, int O_pad, int I, int kh, int kw, int k_pad,
                            void* stream) {
  GDF_LAUNCH(launch_pack_conv_weight(static_cast<const float*>(w), static_cast<bf16*>(out), O, O_pad, I, kh, kw, k_pad,
                                     static_cast<cudaStream_t>(stream)));
}

int64_t gdf_op_groupnorm_workspace_floats(int B, int G) { return (int64_t)gn_workspace_floats(B, G); }

int gdf_op_groupnorm(const void* x, void* y, const void* gamma, const void* beta, int B, int HW, int C, int G,
                     float eps, int silu, void* workspace, void* stream) {
  GDF_LAUNCH(launch_groupnorm(static_cast<const bf16*>(x), static_cast<bf16*>(y), static_cast<const float*>(gamma),
                              static_cast<const float*>(beta), B, HW, C, G, eps, silu != 0,
                              static_cast<float*>(workspace), static_cast<cudaStream_t>(stream)));
}

int gdf_op_layernorm(const void* x, void* y, const void* gamma, const void* beta, int64_t M, int C, float eps,
                     const void* mod_scale, const void* mod_shift, int rows_per_batch, void* stream) {
  GDF_LAUNCH(launch_layernorm(static_cast<const bf16*>(x), static_cast<bf16*>(y), static_cast<const float*>(gamma),
                              static_cast<const float*>(beta), M, C, eps, static_cast<const float*>(mod_scale),
                              static_cast<const float*>(mod_shift), rows_per_batch,
                              static_cast<cudaStream_t>(stream)));
}

int gdf_op_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, int B,
                     int heads, int Nq, int Nk, int head_dim, float scale, int v_f16, void* stream) {
  if (head_dim != 64) {
    if (v_f16) {
      if (!attention_tc_supports(head_dim))
        return fail(GDF_ERR_UNSUPPORTED, "gdf_op_attention: fp16 V needs head_dim 40 / 64 / 72 / 80 / 128");
      return launch_attention_tc(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
                                 static_cast<const bf16*>(v), ldv, static_cast<bf16*>(o), ldo, B, heads, Nq, Nk, head_dim,
                                 scale, 1, nullptr, static_cast<cudaStream_t>(stream));
    }
    GDF_LAUNCH(launch_attention_generic(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
                                        static_cast<const bf16*>(v), ldv, static_cast<bf16*>(o), ldo, B, heads, Nq, Nk,
                                        head_dim, scale, static_cast<cudaStream_t>(stream)));
  }
  GDF_LAUNCH(launch_attention64(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
                                static_cast<const bf16*>(v), ldv, static_cast<bf16*>(o), ldo, B, heads, Nq, Nk, scale,
                                v_f16, static_cast<cudaStream_t>(stream)));
}

int gdf_op_attention_bias(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                          int B, int heads, int Nq, int Nk, int head_dim, float scale, const void* key_bias,
                          void* stream) {
  GDF_LAUNCH(launch_attention_generic(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk,
                                      static_cast<const bf16*>(v), ldv, static_cast<bf16*>(o), ldo, B, heads, Nq, Nk,
                                      head_dim, scale, static_cast<cudaStream_t>(stream),
                                      static_cast<const float*>(key_bias)));
}

/* Debug: the next attention launches record a timeline of CTA 0 (role events stamped with the SM clock) into `buf`
 * (u64[cap], zeroed by the caller; [0] = number of events). NULL turns it off. tools/attn_trace.py decodes it. */
int gdf_debug_attention_trace(void* buf_dev, int cap) {
  attention_tc_set_trace(buf_dev, cap);
  return GDF_OK;
}

int gdf_op_softmax_rows(void* s, int64_t rows, int cols, int ld, void* stream) {
  GDF_LAUNCH(launch_softmax_rows(static_cast<bf16*>(s), rows, cols, ld, static_cast<cudaStream_t>(stream)));
}

int gdf_op_upsample_nearest2x(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  GDF_LAUNCH(launch_upsample_nearest2x(static_cast<const bf16*>(x), static_cast<bf16*>(y), B, H, W, C,
                                       static_cast<cudaStream_t>(stream)));
}

int gdf_op_im2col_small(const void* src_f32, const void* src_bf16, void* a, int B, int H, int W, int Cin,
                        void* stream) {
  GDF_LAUNCH(launch_im2col_small(static_cast<const float*>(src_f32), static_cast<const bf16*>(src_bf16),
                                 static_cast<bf16*>(a), B, H, W, Cin, static_cast<cudaStream_t>(stream)));
}

int gdf_op_qsample(const void* moments, const void* eps_vae, const void* eps_q, float scaling_factor, float sqrt_ab,
                   float sqrt_1m_ab, float input_scale, void* latent_nhwc, void* cap_unet_in, void* latents_nchw,
                   int B, int HW, void* stream) {
  GDF_LAUNCH(launch_qsample(static_cast<const float*>(moments), static_cast<const float*>(eps_vae),
                            static_cast<const float*>(eps_q), scaling_factor, 0.f, sqrt_ab, sqrt_1m_ab, input_scale,
                            static_cast<bf16*>(latent_nhwc), static_cast<__half*>(cap_unet_in),
                            static_cast<float*>(latents_nchw), B, HW, 4, static_cast<cudaStream_t>(stream)));
}

int gdf_op_cast_f32_to_bf16(const void* x, void* y, int64_t n, void* stream) {
  GDF_LAUNCH(launch_cast_f32_to_bf16(static_cast<const float*>(x), static_cast<bf16*>(y), n,
                                     static_cast<cudaStream_t>(stream)));
}

int gdf_op_resize_concat(const gdf_resize_src* srcs, int n_src, int B, int OH, int OW, int Ctot, void* out_nhwc,
                         void* out_nchw, void* sumsq, void* stream) {
  if (n_src > 1024) return fail(GDF_ERR_INVALID, "gdf_op_resize_concat: too many sources (%d > 1024)", n_src);
  std::vector<ResizeSrc> rs(n_src > 0 ? n_src : 1);
  for (int i = 0; i < n_src; ++i) {
    rs[i].ptr = static_cast<const __half*>(srcs[i].ptr_dev);
    rs[i].h = srcs[i].h;
    rs[i].w = srcs[i].w;
    rs[i].C = srcs[i].C;
    rs[i].c_off = srcs[i].c_off;
  }
  GDF_LAUNCH(launch_resize_concat(rs.data(), n_src, B, OH, OW, Ctot, static_cast<__half*>(out_nhwc),
                                  static_cast<__half*>(out_nchw), static_cast<float*>(sumsq),
                                  static_cast<cudaStream_t>(stream)));
}

int gdf_op_avgpool_nhwc(const void* x, void* y, int B, int H, int W, int C, int OH, int OW, void* stream) {
  GDF_LAUNCH(launch_adaptive_avgpool_nhwc(static_cast<const __half*>(x), static_cast<__half*>(y), B, H, W, C, OH, OW,
                                          static_cast<cudaStream_t>(stream)));
}

int64_t gdf_correspond_workspace_floats(int n, int hw, int C) { return (int64_t)corr_workspace_floats(n, hw, C); }

int gdf_correspond(const void* stack_src, const void* stack_tgt, int C, int hw, int load_hw, const void* query_yx,
                   int n, void* idx_out, void* workspace, void* stream) {
  return launch_correspond(static_cast<const __half*>(stack_src), static_cast<const __half*>(stack_tgt), C, hw, load_hw,
                           static_cast<const int*>(query_yx), n, static_cast<long long*>(idx_out),
                           static_cast<float*>(workspace), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
