/*
 * gdf.h - C ABI of the B200-native feature-extraction hot path (libgdf_b200.so).
 *
 * The reference (Darkbblue/generic-diffusion-feature) is pure Python and has no FFI; its "operator API" for
 * this path is the Python surface feature/diffusion_feature.py:26-40 (ctor), :222-235 (extract) plus the
 * internal seam feature/components/feature_extractor.py:83-89 (FeatureGatherer.gather -> FeatureStore.store)
 * and feature/components/models.py:10 (get_diffusion_model). This header is what a maintainer binds (ctypes,
 * see INTEGRATION.md) to replace that seam: the denoiser / VAE forward of the diffusers modules and every
 * gather() call become gdf_encode_noise + gdf_denoise_capture writing straight into a caller-owned arena.
 *
 * Conventions
 *   - every pointer named *_dev is a device pointer owned by the caller (torch allocator); nothing is copied
 *     to the host; all work is enqueued on the passed cudaStream_t (void* here to stay C-only);
 *   - return value 0 = OK, negative = error (GDF_ERR_*); gdf_last_error() returns the message of the last
 *     failure on the calling thread;
 *   - a handle is thread-compatible: one thread at a time per handle, any number of handles per process
 *     (the reference runs one extractor per GPU from Python threads, aggregation_network.py:86-93);
 *   - activations are bf16 NHWC / token-major inside the library; captured features are fp16, laid out
 *     [B, h*w, C] (token-major) in the arena. FeatureStore.store casts every map to fp16
 *     (feature_extractor.py:59-60); ViT maps already come back token-major-strided from the reference
 *     (einops view, :46-48), conv maps are exposed to Python as a permuted (B,C,h,w) view.
 */
#ifndef GDF_H_
#define GDF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GDF_OK 0
#define GDF_ERR_INVALID (-1)
#define GDF_ERR_CUDA (-2)
#define GDF_ERR_UNSUPPORTED (-3)
#define GDF_ERR_MISSING_WEIGHT (-4)
#define GDF_ERR_SHAPE (-5)

/* 3: gdf_epilogue grew (act 4 = ReLU, out_f16_from < 0, res_f16, k_split_*); gdf_plan_decoder / gdf_decode_latents */
#define GDF_ABI_VERSION 3

typedef struct gdf_handle_s* gdf_handle;

const char* gdf_last_error(void);
int gdf_abi_version(void);

/* ------------------------------------------------------------------------------------------------ model level */

/* Architecture description of the denoiser + VAE encoder; mirrors the diffusers config keys the reference
 * relies on (unet/unet_2d_condition.py:171-484 ctor arguments). */
#define GDF_MAX_LEVELS 4
typedef struct gdf_unet_arch {
  int in_channels;                    /* 4 */
  int out_channels;                   /* 4 */
  int num_levels;                     /* len(block_out_channels) */
  int block_out_channels[GDF_MAX_LEVELS];
  int layers_per_block;               /* 2 */
  int down_has_attn[GDF_MAX_LEVELS];  /* CrossAttnDownBlock2D (1) vs DownBlock2D (0) */
  int up_has_attn[GDF_MAX_LEVELS];    /* CrossAttnUpBlock2D vs UpBlock2D, in up_blocks order */
  int transformer_depth[GDF_MAX_LEVELS]; /* transformer_layers_per_block, down order (mid uses the last) */
  int num_heads[GDF_MAX_LEVELS];      /* attention heads per level, down order */
  int cross_attention_dim;            /* 768 / 1024 / 2048 */
  int use_linear_projection;          /* 0: 1x1 conv proj_in/out (SD-1.5), 1: Linear (SD-2.1 / SDXL) */
  int addition_time_embed_dim;        /* 256 for SDXL text_time conditioning, 0 = none */
  int projection_class_embeddings_input_dim; /* 2816 for SDXL */
  int norm_num_groups;                /* 32 */
  float norm_eps;                     /* 1e-5 */
} gdf_unet_arch;

typedef struct gdf_vae_arch {
  int in_channels;                    /* 3 */
  int latent_channels;                /* 4 */
  int num_levels;                     /* 4 */
  int block_out_channels[GDF_MAX_LEVELS]; /* 128,256,512,512 */
  int layers_per_block;               /* 2 */
  int norm_num_groups;                /* 32 */
  float norm_eps;                     /* 1e-6 */
  float scaling_factor;               /* 0.18215 / 0.13025 / 0.3611 (Flux) */
  float shift_factor;                 /* 0 / 0.1159 (Flux): z = (sample - shift_factor) * scaling_factor */
} gdf_vae_arch;

/* PixArt-style DiT denoiser (models.py:71-117 'pixart-sigma' / 'pixart-sigma-512' / 'pixart-alpha'; [diffusers
 * PixArtTransformer2DModel config, un-vendored]; block arithmetic = the reference's vendored
 * BasicTransformerBlock with norm_type 'ada_norm_single', feature/diffusers/models/attention.py:469-592). */
typedef struct gdf_dit_arch {
  int in_channels;                    /* 4 */
  int out_channels;                   /* 8 (learned sigma) */
  int patch_size;                     /* 2 */
  int num_layers;                     /* 28 */
  int num_heads;                      /* 16 */
  int head_dim;                       /* 72 */
  int caption_channels;               /* 4096 (T5) */
  float norm_eps;                     /* 1e-6 */
} gdf_dit_arch;

/* replaces get_diffusion_model (feature/components/models.py:10): the handle owns packed bf16 weights */
int gdf_create(const gdf_unet_arch* unet, const gdf_vae_arch* vae, int device, gdf_handle* out);
/* same, DiT denoiser ("transformer.*" parameter names); gdf_load_weights / gdf_finalize_weights / gdf_plan /
 * gdf_encode_noise / gdf_encode_latents are shared, the forward is gdf_denoise_capture_dit. Feature ids:
 * vit-block{i}-{self-q,self-k,self-v,cross-q,ffn-inner,out} (feature_extractor.py:259-286). */
int gdf_create_dit(const gdf_dit_arch* dit, const gdf_vae_arch* vae, int device, gdf_handle* out);

/* Flux MMDiT denoiser (models.py:150-172 'flux' = FLUX.1-dev; FluxTransformer2DModel defaults at
 * feature/diffusers/models/transformers/transformer_flux.py:253-266; block arithmetic = the reference's vendored
 * FluxTransformerBlock :116-226, FluxSingleTransformerBlock :46-112 and FluxAttnProcessor2_0,
 * attention_processor.py:2259-2362). */
typedef struct gdf_flux_arch {
  int in_channels;                    /* 64 = 16 latent channels x (2 x 2) packed pixels */
  int num_layers;                     /* 19 double-stream (image + text) blocks */
  int num_single_layers;              /* 38 single-stream blocks */
  int num_heads;                      /* 24 */
  int head_dim;                       /* 128 */
  int joint_attention_dim;            /* 4096 (T5) */
  int pooled_projection_dim;          /* 768 (CLIP pooled) */
  int guidance_embeds;                /* 1 for FLUX.1-dev */
} gdf_flux_arch;
/* "transformer.*" parameter names; shared calls as for gdf_create_dit, the forward is gdf_denoise_capture_flux.
 * Feature ids (feature_extractor.py:98-123): vit-block{i}-{q,k,v,attn-out,norm-out,ffn-inner,out} for the double
 * blocks i < num_layers, vit-block{i}-{q,k,v,attn-out,out} for the single blocks numbered after them; image
 * tokens only, (B, 3072, h/16, w/16). */
int gdf_create_flux(const gdf_flux_arch* flux, const gdf_vae_arch* vae, int device, gdf_handle* out);
int gdf_destroy(gdf_handle h);

/* Weights by diffusers parameter name ("unet.down_blocks.0.resnets.0.conv1.weight", "vae.encoder.conv_in.weight",
 * ...), fp32 device tensors in their PyTorch layouts (Linear [out,in], Conv2d OIHW). Packed to bf16 on device.
 * Call repeatedly, then gdf_finalize_weights once (checks that every parameter the architecture needs arrived). */
int gdf_load_weights(gdf_handle h, const char* const* names, const void* const* ptrs_dev, const int64_t* shapes,
                     const int* ranks, int n, void* stream);
/* Packs the weights (bf16, K-major; conv taps; fused / folded matrices), checks every extent against the architecture
 * (GDF_ERR_SHAPE / GDF_ERR_MISSING_WEIGHT name the offending parameter) and releases the fp32 originals of the
 * matrices (GDF_KEEP_FP32_WEIGHTS=1 keeps them). Loading weights again after this drops the plan and every packed
 * tensor: finalise and plan again. */
int gdf_finalize_weights(gdf_handle h, void* stream);

/* Feature plan: replaces prepare_feature_extractor (feature/components/feature_extractor.py:92-288).
 * ids are the reference's feature ids (feature/configs/*.json keys). For each accepted id the library returns
 * the arena slot: byte offset, channels, height, width. cross-k / cross-v are rejected exactly like
 * FeatureStore.store drops them (feature_extractor.py:38-39): offset = -1. Unknown ids -> GDF_ERR_INVALID.
 * Slot layout: fp16 token-major [B, height*width, channels], except
 *   "<block>-self-map" / "<block>-cross-map" (UNet families; the reference's AttnStoreProcessor,
 *       feature/components/attention.py:241-244): fp16 [B, heads, Nq, Nk] with channels = heads, height = Nq,
 *       width = Nk; requesting one switches that attention module to the probability-materialising kernel;
 *   "#attnmean:<block>-self" / "#attnmean:<block>-cross" (internal, not a reference id): the head mean of the same
 *       probabilities, fp16 [B, Nq, Nk] (channels = 1) - what the reference's AttentionStore receives
 *       (attention.py:241-242); the host aggregates these into the `attn` feature (diffusion_feature.py:488-500). */
typedef struct gdf_slot {
  int64_t offset_bytes;               /* -1: id accepted by the grammar but never stored (cross-k/v) */
  int channels, height, width;
  int order;                          /* execution order index (the reference's dict insertion order) */
} gdf_slot;
int gdf_plan(gdf_handle h, const char* const* feature_ids, int n_ids, int batch, int img_size, gdf_slot* slots_out,
             int64_t* arena_bytes_out);
/* A handle holds ONE compiled plan. Every successful gdf_plan bumps this counter; callers that share a handle (two
 * extractors on one pipe, diffusion_feature.py:46-47 `external_model`) compare it with the value they saw after their
 * own gdf_plan and plan again when it moved. 0 = no plan. */
uint64_t gdf_plan_generation(gdf_handle h);
/* Encoder-hidden-state length the next plan is built for (default 77, CLIP). Invalidates the current plan. */
int gdf_set_ctx_len(gdf_handle h, int ctx_len);
/* Introspection of the current plan: kernels launched per (encode + denoise) pass, internal workspace bytes. */
int gdf_num_launches(gdf_handle h);
int64_t gdf_workspace_bytes(gdf_handle h);
/* Profiling pass (bench.py roofline): while enabled, every kernel launch of encode/denoise is bracketed by a
 * CUDA-event pair on the launching stream; device ms / algorithmic FLOPs / launches accumulate per kernel kind
 * (0 tcgen05 GEMM + implicit-GEMM conv, 1 attention, 2 GroupNorm, 3 LayerNorm, 4 other). Arrays of 5. */
int gdf_profile(gdf_handle h, int enable);
int gdf_profile_read(gdf_handle h, float* ms_out, double* flops_out, int* launches_out);
/* CSV of the last profiling pass, one row per kernel launch (phase,index,kind,ms,gflop,tflops,label). */
int gdf_profile_dump(gdf_handle h, const char* path);

/* images -> VAE-encoded, noised, scaled latents (replaces pipe.prepare_latents + scheduler.scale_model_input;
 * pipelines/pixart_alpha/pipeline_pixart_sigma.py:598-677, diffusion_feature.py:371-380,405-406).
 *   images_dev : fp32 (B,3,S,S) in [-1,1];  eps_vae_dev / eps_q_dev : fp32 (B,4,S/8,S/8) injected noise
 *   sqrt_alpha_bar / sqrt_one_minus_alpha_bar : q_sample coefficients of the resolved timestep
 *   input_scale : scale_model_input factor (Euler: 1/sqrt(sigma^2+1) applied on x = z + sigma*eps)
 *   latents_out_dev : optional fp32 (B,4,S/8,S/8) copy of the noised latents (before input scaling) */
int gdf_encode_noise(gdf_handle h, const void* images_dev, const void* eps_vae_dev, const void* eps_q_dev,
                     float sqrt_alpha_bar, float sqrt_one_minus_alpha_bar, float input_scale, void* latents_out_dev,
                     void* stream);

/* `vae-out` (feature/diffusion_feature.py:477-485: latents = scheduler.step(noise_pred, t, latents)[0];
 * vae.decode(latents / scaling_factor)). gdf_plan_decoder builds the decoder op list for the current plan's batch and
 * image size (needs the 'vae.decoder.*' / 'vae.post_quant_conv.*' weights, loaded like every other weight; error
 * GDF_ERR_MISSING_WEIGHT otherwise). gdf_decode_latents computes z = c_latent * latents + c_model * model_out - the
 * scheduler step and the 1 / scaling_factor are both linear in the two tensors, the host supplies the coefficients
 * (schedulers.step_coeffs) - and decodes it:
 *   latents_dev / model_out_dev : fp32 (B, latent, S/8, S/8), the latents_out / noise_pred_out of the calls above
 *                                 (model_out_dev may be NULL: plain vae.decode(c_latent * latents))
 *   image_out_dev               : fp32 NHWC (B, S, S, 3) */
int gdf_plan_decoder(gdf_handle h);
int gdf_decode_latents(gdf_handle h, const void* latents_dev, float c_latent, const void* model_out_dev, float c_model,
                       void* image_out_dev, void* stream);

/* Latents supplied directly: the `image.shape[1] == 4` branch of prepare_latents (pipeline_pixart_sigma.py:623-624)
 * - no VAE pass; x_t = a*latents + b*eps_q (eps_q_dev may be NULL = no noise), model input = x_t * input_scale. */
int gdf_encode_latents(gdf_handle h, const void* latents_dev, const void* eps_q_dev, float sqrt_alpha_bar,
                       float sqrt_one_minus_alpha_bar, float input_scale, void* latents_out_dev, void* stream);

/* One denoiser forward with capture (replaces pipe.unet(...) at diffusion_feature.py:446-465 and every
 * feature_gatherer.gather call site listed in SURVEY.md 2.2).
 *   timestep : resolved scheduler timestep (float, e.g. 50.0)
 *   ctx_dev : fp32 (B, ctx_len, cross_attention_dim) encoder hidden states
 *   pooled_dev : fp32 (B, 1280) pooled text embeds or NULL; add_time_ids_dev : fp32 (B, 6) or NULL
 *   arena_dev : caller-owned arena of arena_bytes (checked against the plan: GDF_ERR_SHAPE when too small);
 *   noise_pred_out_dev : optional fp32 (B,4,h,w) */
int gdf_denoise_capture(gdf_handle h, float timestep, const void* ctx_dev, int ctx_len, const void* pooled_dev,
                        const void* add_time_ids_dev, void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev,
                        void* stream);

/* ControlNet conditioning inputs of the UNet forward (feature/diffusers/models/unet/unet_2d_condition.py:1236-1247
 * down_block_additional_residuals, :1261-1275 mid_block_additional_residual; produced in the reference by
 * feature/components/controlnet.py). fp32 NCHW device tensors (B, C_i, s_i, s_i), borrowed until replaced: one per skip
 * tensor in push order (conv_in output, then every resnet / transformer / downsampler output of the down path) and one
 * for the mid-block output. They are added where the reference adds them; captured maps of the down / mid path are
 * taken before the addition like the reference's gather sites. n_down = 0 and mid_dev = NULL clear them.
 * gdf_control_residual_shapes returns the number of skips and fills (channels, side) per skip. */
int gdf_control_residual_shapes(gdf_handle h, int* channels_out, int* sides_out, int max_n);
int gdf_set_control_residuals(gdf_handle h, const void* const* down_dev, int n_down, const void* mid_dev);

/* One DiT forward with capture (replaces pipe.transformer(...) at diffusion_feature.py:467-474).
 *   ctx_dev : fp32 (B, ctx_len, caption_channels) caption embeddings (T5), ctx_len as set by gdf_set_ctx_len
 *   ctx_mask_dev : fp32 (B, ctx_len), 1 = attend, 0 = masked (bias -10000 like the reference), or NULL = all ones
 *   noise_pred_out_dev : optional fp32 (B, out_channels, h, w) un-patchified model output */
int gdf_denoise_capture_dit(gdf_handle h, float timestep, const void* ctx_dev, int ctx_len, const void* ctx_mask_dev,
                            void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev, void* stream);

/* One Flux forward with capture (replaces self.transformer(...) at pipeline_flux_img2img.py:812-822; the reference
 * returns right after it, :841).
 *   sigma : the flow-match sigma of the resolved step = the pipeline's `timestep / 1000` (also the q_sample mix
 *           x_t = sigma * eps + (1 - sigma) * z done by gdf_encode_noise with sqrt_alpha_bar = 1 - sigma,
 *           sqrt_one_minus_alpha_bar = sigma)
 *   guidance : guidance_scale (1.0 in the reference's call, diffusion_feature.py:246-253); ignored without guidance_embeds
 *   ctx_dev : fp32 (B, ctx_len, joint_attention_dim) T5 embeddings; pooled_dev : fp32 (B, pooled_projection_dim)
 *   rope_cos_dev / rope_sin_dev : fp32 (ctx_len + (h/16)*(w/16), head_dim) rotary tables of FluxPosEmbed for
 *           ids = cat(txt_ids, img_ids) (computed by the host, transformer_flux.py:481-482)
 *   noise_pred_out_dev : optional fp32 (B, (h/16)*(w/16), in_channels) packed model output */
int gdf_denoise_capture_flux(gdf_handle h, float sigma, float guidance, const void* ctx_dev, int ctx_len,
                             const void* pooled_dev, const void* rope_cos_dev, const void* rope_sin_dev,
                             void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------ op level
 * Each hot kernel behind a plain entry point: used by the parity tests, by bench.py's roofline probe and by the
 * Python feature-stack / correspondence helpers. */

typedef struct gdf_capture_seg {
  void* ptr_dev;          /* fp16 destination, row pitch ld */
  int col_begin, col_end, ld;
} gdf_capture_seg;

typedef struct gdf_epilogue {
  float alpha;            /* accumulator scale; 0 is treated as 1 */
  int n_out;              /* valid output columns, 0 = all */
  const void* bias_dev;           /* fp32 [N] */
  const void* bias_m_dev;         /* fp32 [M] */
  const void* row_batch_bias_dev; /* fp32 [M/rows_per_batch, N] */
  int rows_per_batch;
  int act;                /* 0 none, 1 GEGLU(erf), 2 GELU-tanh, 3 SiLU, 4 ReLU */
  const void* col_scale_dev;      /* fp32 [M/rows_per_batch, n_out] */
  const void* residual_dev; int ld_res;    /* bf16 */
  float out_scale;        /* 0 is treated as 1 */
  void* out_dev; int ld_out; int64_t out_batch_stride;   /* bf16 */
  int out_f16_from;       /* > 0: columns >= this are written to out as fp16 (V of a fused QKV projection);
                           * < 0: every column of out is fp16 (intermediate of an fp16-operand head) */
  void* out2_dev; int ld_out2;                           /* bf16 */
  void* out_f32_dev; int ld_out_f32;
  void* cap_pre_dev; int ld_cap_pre;                     /* fp16, before residual */
  gdf_capture_seg cap[3]; int num_cap;                   /* fp16, final value */
  /* LayerNorm folded into the consuming projection (attention.py:497,525,566): A holds the un-normalised rows, the
   * weight carries gamma, bias carries beta.W^T: value = rstd[row] * (acc - mean[row] * ln_u[col]) + bias[col] */
  const void* ln_sums_dev;        /* fp32 [M][2] (sum, sum of squares) of the A rows, or NULL */
  const void* ln_u_dev;           /* fp32 [N] row sums of the gamma-folded bf16 weight */
  float ln_eps;
  void* row_sums_dev;             /* fp32 [M][2]: the launch ADDS (sum, sum sq) of its final output rows, or NULL */
  /* GroupNorm statistics of the output for the GroupNorm that consumes it (resnet.py:327,351): the launch ADDS
   * (sum, sum sq) per (image, group); groups of gn_cpg = 4 / 8 / 16 channels, gn_rows_per_img % 128 == 0 */
  void* gn_sums_dev;              /* fp32 [images][gn_groups][2] or NULL */
  int gn_cpg, gn_groups;
  int64_t gn_rows_per_img;
  /* 1: A and W hold fp16 bit patterns instead of bf16 (operands of one 16-bit type: a convolution or projection over an
   * fp16 feature stack - the downstream heads of aggregation_network.py:22,97-99 - with weights from
   * gdf_op_pack_conv_weight_f16) */
  int in_f16;
  /* 1: residual_dev holds fp16 (a captured feature map: the ResBlock heads of
   * segmentation/models/diffusion_segmentor.py:23-44 add their fp16 input back) instead of bf16 */
  int res_f16;
  /* Optional K-split of the last partial wave of gdf_op_linear (batch 1): when the tiles do not fill a whole number of
   * waves of the resident CTA groups, the tiles of the last wave are split along K into pieces that run side by side;
   * partial fp32 accumulators go through k_split_ws_dev and are summed in a fixed order (deterministic). One launch at
   * a time may use a workspace. k_split_cnt_dev must be zero before the first launch (launches leave it zero).
   * Sizes: 74 x 2 x 128 x 256 floats and 74 x 2 x 8 counters are always enough. NULL: whole tiles only. */
  void* k_split_ws_dev; int64_t k_split_ws_floats;
  void* k_split_cnt_dev; int k_split_cnt_len;
} gdf_epilogue;

/* C[M,N] = A[M,K] W[N,K]^T (+ fused epilogue); batch > 1: A/out strided by *_batch_stride elements,
 * W shared if w_batch_stride == 0. (nn.Linear: attention_processor.py:3281-3289,3319; attention.py:1253-1257) */
int gdf_op_linear(const void* a_dev, int64_t M, int K, int lda, const void* w_dev, int N, int ldw,
                  const gdf_epilogue* ep, int batch, int64_t a_batch_stride, int64_t w_batch_stride, int block_n,
                  void* stream);
/* 3x3 convolution as implicit GEMM over NHWC bf16 (nn.Conv2d: resnet.py:341,366; downsampling.py:147).
 * w_packed_dev: bf16 [N][9*Cin], k = (ky*3+kx)*Cin + c (gdf_op_pack_conv_weight). */
int gdf_op_conv3x3(const void* x_dev, int B, int Hin, int Win, int Cin, const void* w_packed_dev, int N, int stride,
                   int pad_lo, const gdf_epilogue* ep, int block_n, void* stream);
int gdf_op_pack_conv_weight(const void* w_oihw_f32_dev, void* out_bf16_dev, int O, int O_pad, int I, int kh, int kw,
                            int k_pad, void* stream);
/* First convolution of the VAE encoder fused from the image ([diffusers Encoder.conv_in], 3 -> N channels, 3x3, pad 1):
 * img fp32 NCHW (B,3,H,W) -> out bf16 NHWC [B*H*W, N] (+ bias), the 27-tap operand is built in shared memory.
 * w_packed: gdf_op_pack_conv_weight(..., k_pad = 64). N in {64, 128}, W % 128 == 0. gn_sums (optional): fp32
 * [B][gn_groups][2], the launch ADDS (sum, sum sq) of its output per (image, group of gn_cpg = 4 / 8 / 16 channels). */
int gdf_op_conv_in(const void* img_nchw_f32_dev, const void* w_packed_dev, const void* bias_dev, void* out_dev, int B,
                   int H, int W, int N, void* gn_sums_dev, int gn_cpg, int gn_groups, void* stream);
int gdf_op_pack_conv_weight_f16(const void* w_oihw_f32_dev, void* out_f16_dev, int O, int O_pad, int I, int kh, int kw,
                                int k_pad, void* stream);
/* GroupNorm(+SiLU) NHWC bf16 (resnet.py:327-328). workspace_dev: fp32, gdf_op_groupnorm_workspace_floats(B,G). */
int64_t gdf_op_groupnorm_workspace_floats(int B, int G);
int gdf_op_groupnorm(const void* x_dev, void* y_dev, const void* gamma_dev, const void* beta_dev, int B, int HW,
                     int C, int G, float eps, int silu, void* workspace_dev, void* stream);
/* LayerNorm (+ AdaLN-single modulation) (attention.py:498-503,539,565) */
int gdf_op_layernorm(const void* x_dev, void* y_dev, const void* gamma_dev, const void* beta_dev, int64_t M, int C,
                     float eps, const void* mod_scale_dev, const void* mod_shift_dev, int rows_per_batch,
                     void* stream);
/* softmax(QK^T*scale)V, head_dim 64 (attention_processor.py:3311-3313). q, k bf16; v bf16 (v_f16 = 0, mma.sync
 * kernel) or fp16 bit patterns (v_f16 = 1, tcgen05/TMEM kernel, needs Nk >= 128). */
int gdf_op_attention(const void* q_dev, int ldq, const void* k_dev, int ldk, const void* v_dev, int ldv, void* o_dev,
                     int ldo, int B, int heads, int Nq, int Nk, int head_dim, float scale, int v_f16, void* stream);
/* any head_dim (multiple of 8, <= 160), bf16 v, optional additive key bias fp32 (B, Nk): PixArt masked cross-attention */
int gdf_op_attention_bias(const void* q_dev, int ldq, const void* k_dev, int ldk, const void* v_dev, int ldv,
                          void* o_dev, int ldo, int B, int heads, int Nq, int Nk, int head_dim, float scale,
                          const void* key_bias_dev, void* stream);
int gdf_op_softmax_rows(void* s_dev, int64_t rows, int cols, int ld, void* stream);
/* Debug aid: CTA 0 of the following attention launches records (event id << 40 | SM clock) into buf_dev (u64[cap],
 * zero it first; [0] = event count). NULL turns tracing off. Event ids: tools/attn_trace.py. */
int gdf_debug_attention_trace(void* buf_dev, int cap);
int gdf_op_upsample_nearest2x(const void* x_dev, void* y_dev, int B, int H, int W, int C, void* stream);
int gdf_op_im2col_small(const void* src_nchw_f32_dev, const void* src_nhwc_bf16_dev, void* a_dev, int B, int H,
                        int W, int Cin, void* stream);
int gdf_op_qsample(const void* moments_dev, const void* eps_vae_dev, const void* eps_q_dev, float scaling_factor,
                   float sqrt_ab, float sqrt_1m_ab, float input_scale, void* latent_nhwc_dev, void* cap_unet_in_dev,
                   void* latents_nchw_f32_dev, int B, int HW, void* stream);
int gdf_op_cast_f32_to_bf16(const void* x_dev, void* y_dev, int64_t n, void* stream);

/* Feature stack (aggregation_network.py:62-66): bilinear resize of n_src captured maps to (OH, OW) + concat. */
typedef struct gdf_resize_src {
  const void* ptr_dev;    /* fp16 [B, h*w, C] */
  int h, w, C, c_off;
} gdf_resize_src;
int gdf_op_resize_concat(const gdf_resize_src* srcs, int n_src, int B, int OH, int OW, int Ctot, void* out_nhwc_dev,
                         void* out_nchw_dev, void* sumsq_dev, void* stream);

/* feature_resize (FeatureStore.store, feature/components/feature_extractor.py:51-53): F.adaptive_avg_pool2d of one
 * captured map, fp16 NHWC [B, H*W, C] -> [B, OH*OW, C] (OH = H // resize_ratio). */
int gdf_op_avgpool_nhwc(const void* x_dev, void* y_dev, int B, int H, int W, int C, int OH, int OW, void* stream);

/* Correspondence (correspondence/correspondence/correspondence_utils.py:113-146 find_nn_source_correspondences):
 * for n query points (int32 (n,2) [y,x], already rounded/clipped like points_to_idxs :140-146) on the
 * load_hw x load_hw grid of the source image, the flat index (int64, y*load_hw + x) of the most cosine-similar
 * position of the target image. stacks: fp16 NHWC [hw*hw, C] (gdf_op_resize_concat, one image each).
 * Evaluated in the exact low-resolution form (similarity GEMM n x hw^2 x C on the tensor cores + interpolated
 * arg-max); workspace: fp32, gdf_correspond_workspace_floats(n, hw, C). */
int64_t gdf_correspond_workspace_floats(int n, int hw, int C);
int gdf_correspond(const void* stack_src_dev, const void* stack_tgt_dev, int C, int hw, int load_hw,
                   const void* query_yx_dev, int n, void* idx_out_dev, void* workspace_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GDF_H_ */
