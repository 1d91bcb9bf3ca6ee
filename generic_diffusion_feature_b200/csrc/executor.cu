// Model-level C ABI (include/gdf.h): weights, feature plan, VAE-encode + q_sample, denoiser forward with
// capture. The runtime is native: gdf_plan walks the architecture once, allocates every activation buffer,
// builds every TMA descriptor / GEMM launch and records a flat op list; gdf_encode_noise and
// gdf_denoise_capture replay that list on the caller's stream (no Python in the loop, no host sync).
//
// Reference control flow being replaced (file:line under /root/reference/feature):
//   diffusion_feature.py:371-380,405-406   prepare_latents + scale_model_input      -> gdf_encode_noise
//   diffusion_feature.py:446-465            pipe.unet(...)                            -> gdf_denoise_capture
//   diffusers/models/unet/unet_2d_condition.py:1040-1319 (forward), resnet.py:320-379, transformers/
//   transformer_2d.py:403-530, attention.py:469-592,1249-1258, attention_processor.py:3244-3331,
//   downsampling.py:132-152, upsampling.py:142-195                                  -> op list below
//   components/feature_extractor.py:31-76,92-288 (store + id grammar)                -> capture slots
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ops.h"
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3 (ranges behind GDF_NVTX=1)

namespace gdf {

struct RawW {
  float* ptr = nullptr;          // fp32 copy on the device; null once released (see gdf_finalize_weights)
  std::vector<int64_t> shape;
  int64_t numel = 0;
  bool keep = false;             // read directly at run time / plan time (biases, norm affines, conditioning MLPs)
  bool packed_from = false;      // a bf16 / conv / folded packing was made from it (the only tensors ever released)
};

struct RunCtx {
  cudaStream_t stream;
  char* arena = nullptr;
  const float* images = nullptr;
  const float* eps_vae = nullptr;
  const float* eps_q = nullptr;
  float qa = 1.f, qb = 0.f, qs = 1.f;
  float* latents_out = nullptr;
  float* noise_pred_out = nullptr;
  // gdf_decode_latents (vae-out): z = dec_c_latent * dec_latents + dec_c_model * dec_model_out, image out fp32 NHWC
  const float* dec_latents = nullptr;
  const float* dec_model_out = nullptr;
  float dec_c_latent = 1.f, dec_c_model = 0.f;
  float* dec_image_out = nullptr;
};
typedef std::function<int(const RunCtx&)> Op;

enum OpKind : int { kKindGemm = 0, kKindAttention = 1, kKindGroupNorm = 2, kKindLayerNorm = 3, kKindOther = 4,
                    kNumKinds = 5 };
struct OpList {   // flat program: closures + bookkeeping for the profiling pass
  std::vector<Op> fns;
  std::vector<int> kinds;
  std::vector<double> flops;
  std::vector<std::string> labels;
  std::vector<std::string> scopes;   // block the op belongs to ("down-level1-repeat0-vit-block0", "vae", ...): NVTX ranges
  std::string scope;                 // current block, set by the emit_* functions; sticks until the next one
  std::vector<float> last_ms;
  int cur_kind = kKindOther;
  double cur_flops = 0.0;
  std::string cur_label;
  void push_back(Op f) {
    fns.push_back(std::move(f));
    kinds.push_back(cur_kind);
    flops.push_back(cur_flops);
    labels.push_back(cur_label);
    scopes.push_back(scope);
    last_ms.push_back(0.f);
    cur_kind = kKindOther;
    cur_flops = 0.0;
    cur_label.clear();
  }
  void tag(int kind, double fl, const std::string& label = std::string()) {
    cur_kind = kind;
    cur_flops = fl;
    cur_label = label;
  }
  void clear() {
    fns.clear();
    kinds.clear();
    flops.clear();
    labels.clear();
    scopes.clear();
    scope.clear();
    last_ms.clear();
  }
  size_t size() const { return fns.size(); }
};

#define OP_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return fail(GDF_ERR_CUDA, "%s -> %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// ----------------------------------------------------------------------------------------- small kernels
__global__ void pack_rows_kernel(const float* __restrict__ src, int K, const int* __restrict__ row_idx, int R_out,
                                 bf16* __restrict__ dst) {
  const long long total = (long long)R_out * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i % K);
    const int sr = row_idx ? row_idx[r] : r;
    dst[i] = __float2bfloat16_rn(sr >= 0 ? src[(long long)sr * K + k] : 0.f);
  }
}
// LayerNorm folded into a projection: dst[r, k] = bf16(src[idx[r], k] * gamma[k])
__global__ void pack_rows_scaled_kernel(const float* __restrict__ src, int K, const int* __restrict__ row_idx, int R_out,
                                        const float* __restrict__ gamma, bf16* __restrict__ dst) {
  const long long total = (long long)R_out * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), k = (int)(i % K);
    const int sr = row_idx ? row_idx[r] : r;
    dst[i] = __float2bfloat16_rn(sr >= 0 ? src[(long long)sr * K + k] * gamma[k] : 0.f);
  }
}
// ... and its per-row constants: u[r] = sum_k dst[r, k] (the bf16 values the tensor core multiplies),
// c[r] = bias[idx[r]] + sum_k src[idx[r], k] * beta[k]. One warp per row.
__global__ void ln_fold_consts_kernel(const float* __restrict__ src, const bf16* __restrict__ dst, int K,
                                      const int* __restrict__ row_idx, int R_out, const float* __restrict__ beta,
                                      const float* __restrict__ bias, float* __restrict__ u, float* __restrict__ c) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R_out) return;
  const int sr = row_idx ? row_idx[r] : r;
  float su = 0.f, sc = 0.f;
  for (int k = lane; k < K; k += 32) {
    su += __bfloat162float(dst[(long long)r * K + k]);
    if (sr >= 0) sc += src[(long long)sr * K + k] * beta[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if (lane == 0) {
    u[r] = su;
    c[r] = sc + ((bias && sr >= 0) ? bias[sr] : 0.f);
  }
}
__global__ void gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, int n,
                                  float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (idx[i] >= 0) ? src[idx[i]] : 0.f;
}
__global__ void fill_f32_kernel(float* dst, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = v;
}
__global__ void add_f32_kernel(float* dst, const float* a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += a[i];
}
// latents given directly (prepare_latents, image.shape[1] == 4 branch, pipeline_pixart_sigma.py:623-624):
// x_t = a*z + b*eps, model input = x_t * s; NCHW fp32 in, NHWC bf16 out
__global__ void latents_qsample_kernel(const float* __restrict__ z, const float* __restrict__ eps, float a, float b,
                                       float s, bf16* __restrict__ latent_nhwc, float* __restrict__ latents_out, int B,
                                       int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long r = i / HW;
    const int c = (int)(r % C), bb = (int)(r / C);
    const float xt = a * z[i] + b * (eps ? eps[i] : 0.f);
    if (latents_out) latents_out[i] = xt;
    latent_nhwc[((long long)bb * HW + p) * C + c] = __float2bfloat16_rn(xt * s);
  }
}
// ControlNet residual (fp32 NCHW (B, C, HW)) added into a bf16 NHWC tensor / column slice: dst[(b*HW + p)*ld + c] += r
// (unet_2d_condition.py:1236-1247 down_block_additional_residuals, :1261-1275 mid_block_additional_residual)
__global__ void add_residual_nchw_kernel(bf16* __restrict__ dst, int ld, const float* __restrict__ res, int B, int HW,
                                         int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long r = i / C;
    const int p = (int)(r % HW), b = (int)(r / HW);
    bf16* d = dst + ((long long)b * HW + p) * ld + c;
    *d = __float2bfloat16_rn(__bfloat162float(*d) + res[((long long)b * C + c) * HW + p]);
  }
}
// moments/noise-pred style NHWC bf16/f32 [B, HW, C] -> NCHW fp32 (B, C, HW)
__global__ void nhwc_bf16_to_nchw_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW);
    const long long r = i / HW;
    const int c = (int)(r % C), b = (int)(r / C);
    y[i] = __bfloat162float(x[((long long)b * HW + p) * C + c]);
  }
}

// ----------------------------------------------------------------------------------------- buffer pool
class BufPool {
 public:
  ~BufPool() { clear(); }
  void* acquire(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    auto it = free_.lower_bound(bytes);
    if (it != free_.end() && it->first <= bytes * 2 + (1 << 20)) {
      void* p = it->second;
      size_of_[p] = it->first;
      free_.erase(it);
      return p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    // GDF_POISON_WORKSPACE=1 (debug): every new workspace buffer starts as 0xFF bytes (bf16 / fp16 / fp32 NaN), so a kernel
    // that reads a buffer region nobody wrote in this forward shows up as NaN instead of as a silent run-to-run difference
    static const bool poison = [] { const char* e = getenv("GDF_POISON_WORKSPACE"); return e && e[0] == '1'; }();
    if (poison) cudaMemset(p, 0xFF, bytes);
    all_.push_back(p);
    size_of_[p] = bytes;
    total_ += bytes;
    return p;
  }
  void release(void* p) {
    if (!p) return;
    auto it = size_of_.find(p);
    if (it == size_of_.end()) return;
    free_.insert({it->second, p});
  }
  void clear() {
    for (void* p : all_) cudaFree(p);
    all_.clear();
    free_.clear();
    size_of_.clear();
    total_ = 0;
  }
  size_t total() const { return total_; }

 private:
  std::multimap<size_t, void*> free_;
  std::unordered_map<void*, size_t> size_of_;
  std::vector<void*> all_;
  size_t total_ = 0;
};

struct Site {  // one capture call site of the reference, in execution order
  std::string id;
  int C, H, W;
};

}  // namespace gdf

using namespace gdf;

struct gdf_handle_s {
  gdf_unet_arch ua;
  gdf_vae_arch va;
  gdf_dit_arch da;
  gdf_flux_arch fa;
  bool is_dit = false;
  bool is_flux = false;
  const float* rope_cos = nullptr;   // [ctx_len + N_img, head_dim] fp32, borrowed per call (Flux)
  const float* rope_sin = nullptr;
  float* g_dev = nullptr;            // [B] guidance * 1000 (Flux)
  float* key_bias = nullptr;     // [B, ctx_len] additive cross-attention bias (DiT), valid when has_key_bias
  bool has_key_bias = false;
  int device = 0;
  std::unordered_map<std::string, RawW> raw;
  std::unordered_map<std::string, void*> packed;
  std::vector<void*> owned;
  std::unordered_map<const void*, std::pair<int64_t, int64_t>> mat_dims;   // packed matrix -> (rows, cols)
  std::unordered_map<const void*, int64_t> vec_len;                        // fp32 vector -> elements
  uint64_t plan_generation = 0;      // bumped by every successful gdf_plan (gdf_plan_generation)
  bool finalized = false;
  // plan
  bool planned = false;
  int B = 0, img = 0, L = 0, ctx_len = 77;
  OpList vae_ops, unet_ops;
  OpList dec_ops;                    // VAE decoder of the `vae-out` path (gdf_plan_decoder), empty unless requested
  bool profile = false;
  float prof_ms[kNumKinds] = {0, 0, 0, 0, 0};
  double prof_flops[kNumKinds] = {0, 0, 0, 0, 0};
  int prof_launches[kNumKinds] = {0, 0, 0, 0, 0};
  BufPool pool;
  std::unordered_map<std::string, int> requested;  // id -> index in the caller's list
  std::vector<gdf_slot> slots;
  std::vector<Site> sites;
  int64_t arena_bytes = 0;
  // run-time staging buffers (device)
  float* t_dev = nullptr;        // [B]
  bf16* ctx_bf16 = nullptr;      // [B*ctx_len, ctx_dim]
  float* add_in = nullptr;       // [B, add_in_dim]
  float* time_ids_dev = nullptr; // [B*6] pointer set per call (borrowed)
  const float* ctx_f32 = nullptr;
  const float* pooled = nullptr;
  bf16* latent_nhwc = nullptr;   // [B, L*L, 4] model input
  int64_t unet_in_cap = -1;
  int gpu_launches = 0;
  // ControlNet residual inputs of the next UNet forward (borrowed device pointers, fp32 NCHW); n_ctrl = 0: none
  std::vector<const float*> ctrl_down;
  const float* ctrl_mid = nullptr;
  int n_skips = 0;                   // skip tensors of the planned UNet (= expected number of down residuals)
  // K-split tail wave of the linear layers (GemmParams::sk_*): one workspace per pipe, used by one launch at a time (the
  // op lists run on one stream); the counters are zeroed here once and left zero by every launch
  float* sk_ws = nullptr;
  unsigned int* sk_cnt = nullptr;
  std::vector<std::pair<int, int>> skip_shapes;   // (channels, side) per skip, push order
};

namespace gdf {

struct Caps {  // arena offsets of the fp16 side outputs of one GEMM (-1 = not requested)
  int64_t pre = -1;
  int64_t post[3] = {-1, -1, -1};
  int c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0};
  int n = 0;
  void add(int64_t off, int col_begin, int col_end) {
    post[n] = off;
    c0[n] = col_begin;
    c1[n] = col_end;
    ++n;
  }
};

// ----------------------------------------------------------------------------------------- builder
class Builder {
 public:
  Builder(gdf_handle_s* h, bool dry) : h(h), dry(dry) {}
  gdf_handle_s* h;
  bool dry;
  int err = 0;
  OpList* ops = nullptr;
  float* gn_ws = nullptr;
  // UNet: the cross-attention K/V of every transformer block depend only on the context, so they come from ONE
  // batched projection at the head of the op list (weights of all blocks row-concatenated) instead of one small
  // M = B * ctx_len GEMM per block; each block reads its [K | V] column slice of kv_all (pitch kv_ld).
  // UNet transformer blocks: LayerNorm folded into the projection that consumes it (GDF_LN_FOLD=0 disables). The
  // producer of the residual stream adds per-row (sum, sum sq) into a slice of ln_slab (zeroed once per forward).
  bool ln_fold = false;
  float* ln_slab = nullptr;
  long long ln_off = 0, ln_total = 0;
  float* ln_rows(long long M) {
    float* p = ln_slab ? ln_slab + ln_off : nullptr;
    ln_off += 2 * M;
    return p;
  }
  // GroupNorm statistics produced by the epilogue of the GEMM that writes the tensor (GDF_GN_FUSE=0 disables):
  // slices of gn_slab, [B][G][2] fp32 each, zeroed by one memset at the head of the op list.
  float* gn_slab = nullptr;
  int gn_used = 0, gn_cap = 0;
  bool gn_fuse = false;
  bool gn_fusable(int C, int G, long long HW) const {
    const int cpg = G > 0 ? C / G : 0;
    return gn_fuse && (cpg == 4 || cpg == 8 || cpg == 16) && C % 64 == 0 && HW % 128 == 0;
  }
  float* gn_slot(int B, int G) {   // null when the slab is exhausted (callers fall back to the statistics kernel)
    if (gn_used >= gn_cap) return nullptr;
    float* p = gn_slab ? gn_slab + (size_t)gn_used * B * G * 2 : nullptr;
    ++gn_used;
    return dry ? reinterpret_cast<float*>(uintptr_t(16)) : p;
  }
  void gn_begin(int B, int G, int max_slots) {   // call once per op list, before the first producer
    const char* ev = getenv("GDF_GN_FUSE");
    gn_fuse = !(ev && ev[0] == '0') && !deterministic_mode();   // epilogue statistics are float atomics
    gn_used = 0;
    gn_cap = gn_fuse ? max_slots : 0;
    gn_slab = gn_fuse ? fbuf((long long)max_slots * B * G * 2) : nullptr;
    if (gn_fuse && !dry) {
      float* slab = gn_slab;
      const size_t bytes = (size_t)max_slots * B * G * 2 * 4;
      ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(cudaMemsetAsync(slab, 0, bytes, rc.stream));
        return 0;
      });
    }
  }
  static void want_gn_stats(Epilogue& e, float* sums, int C, int G, long long HW) {
    if (!sums) return;
    e.gn_sums = sums;
    e.gn_cpg = C / G;
    e.gn_groups = G;
    e.gn_rows_per_img = HW;
  }
  bool kv_batched = false;
  std::vector<std::string> kv_names;
  bf16* kv_all = nullptr;
  int kv_ld = 0, kv_col = 0;

  int set_err(int e) {
    if (!err) err = e;
    return e;
  }
  // ---- weights -------------------------------------------------------------------------------
  const RawW* raw(const std::string& name) {
    auto it = h->raw.find(name);
    if (it == h->raw.end()) {
      set_err(fail(GDF_ERR_MISSING_WEIGHT, "missing weight '%s'", name.c_str()));
      return nullptr;
    }
    return &it->second;
  }
  void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
      set_err(fail(GDF_ERR_CUDA, "cudaMalloc(%zu) failed for weights", bytes));
      return nullptr;
    }
    h->owned.push_back(p);
    return p;
  }
  // fp32 tensor read in place at run time: stays resident after gdf_finalize_weights
  const float* f32(const std::string& name) {
    auto it = h->raw.find(name);
    if (it == h->raw.end()) {
      set_err(fail(GDF_ERR_MISSING_WEIGHT, "missing weight '%s'", name.c_str()));
      return nullptr;
    }
    it->second.keep = true;
    if (!it->second.ptr) {
      set_err(fail(GDF_ERR_INVALID, "weight '%s' was released by gdf_finalize_weights but is read in place by this plan; "
                   "load the weights again (or set GDF_KEEP_FP32_WEIGHTS=1)", name.c_str()));
      return nullptr;
    }
    note_vec(it->second.ptr, it->second.numel);
    return it->second.ptr;
  }
  // fp32 source of a packing step: fails when the tensor was released after finalisation
  const float* src_ptr(const RawW* r, const std::string& name) {
    if (r) const_cast<RawW*>(r)->packed_from = true;
    if (r && !r->ptr) set_err(fail(GDF_ERR_INVALID, "weight '%s' was released by gdf_finalize_weights; load it again "
                                   "before asking for a new packing", name.c_str()));
    return r ? r->ptr : nullptr;
  }
  // ---- shape registry: every packed matrix / fp32 vector handed to an op is remembered with its extents, and the op
  // emitters check them against the extents the ARCHITECTURE asks for (a checkpoint / config mismatch becomes
  // GDF_ERR_SHAPE at finalisation instead of an out-of-bounds read on the device).
  void note_mat(const void* p, int64_t rows, int64_t cols) { if (p) h->mat_dims[p] = {rows, cols}; }
  void note_vec(const void* p, int64_t n) { if (p) h->vec_len[p] = n; }
  bool check_mat(const void* p, int64_t rows_needed, int64_t cols, const char* what) {
    auto it = h->mat_dims.find(p);
    if (it == h->mat_dims.end()) return true;   // activations / slices: not a registered weight
    if (it->second.first < rows_needed || it->second.second != cols) {
      set_err(fail(GDF_ERR_SHAPE, "%s: weight is [%lld, %lld], the architecture needs [>=%lld, %lld]", what,
                   (long long)it->second.first, (long long)it->second.second, (long long)rows_needed, (long long)cols));
      return false;
    }
    return true;
  }
  bool check_vec(const void* p, int64_t n_needed, const char* what) {
    auto it = h->vec_len.find(p);
    if (it == h->vec_len.end()) return true;
    if (it->second < n_needed) {
      set_err(fail(GDF_ERR_SHAPE, "%s: vector holds %lld values, the architecture needs %lld", what,
                   (long long)it->second, (long long)n_needed));
      return false;
    }
    return true;
  }
  // fp32 vector zero-padded to n_pad entries
  const float* f32_pad(const std::string& name, int n_pad) {
    const std::string key = name + "#pad" + std::to_string(n_pad);
    auto it = h->packed.find(key);
    if (it != h->packed.end()) return static_cast<const float*>(it->second);
    const RawW* r = raw(name);
    if (!r || !src_ptr(r, name)) return nullptr;
    if (r->numel > n_pad) {
      set_err(fail(GDF_ERR_SHAPE, "%s holds %lld values, the architecture needs at most %d", name.c_str(),
                   (long long)r->numel, n_pad));
      return nullptr;
    }
    float* p = static_cast<float*>(dev_alloc((size_t)n_pad * 4));
    if (!p) return nullptr;
    cudaMemset(p, 0, (size_t)n_pad * 4);
    cudaMemcpy(p, r->ptr, (size_t)r->numel * 4, cudaMemcpyDeviceToDevice);
    h->packed[key] = p;
    note_vec(p, r->numel);
    return p;
  }
  // rows of [R, K] fp32 matrices stacked and cast to bf16; idx (host) selects/permutes rows of the stack
  const bf16* rows_bf16(const std::string& key, const std::vector<std::string>& names, const std::vector<int>* idx) {
    auto it = h->packed.find(key);
    if (it != h->packed.end()) return static_cast<const bf16*>(it->second);
    int K = -1;
    int64_t R = 0;
    for (auto& n : names) {
      const RawW* r = raw(n);
      if (!r || !src_ptr(r, n)) return nullptr;
      if (r->shape.empty() || r->shape[0] < 1) {
        set_err(fail(GDF_ERR_SHAPE, "%s: not a matrix", n.c_str()));
        return nullptr;
      }
      const int64_t k = r->numel / r->shape[0];
      if (K < 0) K = (int)k;
      if (k != K) {
        set_err(fail(GDF_ERR_SHAPE, "rows_bf16 %s: inner size mismatch", key.c_str()));
        return nullptr;
      }
      R += r->shape[0];
    }
    float* stack = nullptr;
    const float* src = nullptr;
    if (names.size() == 1) {
      src = raw(names[0])->ptr;
    } else {
      cudaMalloc(&stack, (size_t)R * K * 4);
      int64_t off = 0;
      for (auto& n : names) {
        const RawW* r = raw(n);
        cudaMemcpy(stack + off, r->ptr, (size_t)r->numel * 4, cudaMemcpyDeviceToDevice);
        off += r->numel;
      }
      src = stack;
    }
    const int R_out = idx ? (int)idx->size() : (int)R;
    if (idx)
      for (int v : *idx)
        if (v >= R) {
          if (stack) cudaFree(stack);
          set_err(fail(GDF_ERR_SHAPE, "%s: the architecture addresses row %d of a %lld-row weight", key.c_str(), v,
                       (long long)R));
          return nullptr;
        }
    int* idx_dev = nullptr;
    if (idx) {
      cudaMalloc(&idx_dev, idx->size() * 4);
      cudaMemcpy(idx_dev, idx->data(), idx->size() * 4, cudaMemcpyHostToDevice);
    }
    bf16* dst = static_cast<bf16*>(dev_alloc((size_t)R_out * K * 2));
    if (dst) {
      pack_rows_kernel<<<1024, 256>>>(src, K, idx_dev, R_out, dst);
      cudaDeviceSynchronize();
    }
    if (stack) cudaFree(stack);
    if (idx_dev) cudaFree(idx_dev);
    h->packed[key] = dst;
    note_mat(dst, R_out, K);
    return dst;
  }
  // rows_bf16 with LayerNorm (gamma, beta) folded in: weight columns scaled by gamma; u / c per output row (see
  // ln_fold_consts_kernel). bias_name may be empty.
  struct LnW { const bf16* w = nullptr; const float* u = nullptr; const float* c = nullptr; };
  LnW rows_bf16_ln(const std::string& key, const std::vector<std::string>& names, const std::vector<int>* idx,
                   const std::string& norm_prefix, const std::string& bias_name) {
    LnW out;
    auto it = h->packed.find(key);
    if (it != h->packed.end()) {
      out.w = static_cast<const bf16*>(it->second);
      out.u = static_cast<const float*>(h->packed[key + "#u"]);
      out.c = static_cast<const float*>(h->packed[key + "#c"]);
      return out;
    }
    const float* gamma = f32(norm_prefix + ".weight");
    const float* beta = f32(norm_prefix + ".bias");
    const float* bias = bias_name.empty() ? nullptr : f32(bias_name);
    if (!gamma || !beta || (!bias_name.empty() && !bias)) return out;
    int K = -1;
    int64_t R = 0;
    for (auto& n : names) {
      const RawW* r = raw(n);
      if (!r || !src_ptr(r, n)) return out;
      if (r->shape.empty() || r->shape[0] < 1) {
        set_err(fail(GDF_ERR_SHAPE, "%s: not a matrix", n.c_str()));
        return out;
      }
      const int64_t k = r->numel / r->shape[0];
      if (K < 0) K = (int)k;
      if (k != K) {
        set_err(fail(GDF_ERR_SHAPE, "rows_bf16_ln %s: inner size mismatch", key.c_str()));
        return out;
      }
      R += r->shape[0];
    }
    float* stack = nullptr;
    const float* src = nullptr;
    if (names.size() == 1) {
      src = raw(names[0])->ptr;
    } else {
      cudaMalloc(&stack, (size_t)R * K * 4);
      int64_t off = 0;
      for (auto& n : names) {
        const RawW* r = raw(n);
        cudaMemcpy(stack + off, r->ptr, (size_t)r->numel * 4, cudaMemcpyDeviceToDevice);
        off += r->numel;
      }
      src = stack;
    }
    const int R_out = idx ? (int)idx->size() : (int)R;
    if (idx)
      for (int v : *idx)
        if (v >= R) {
          if (stack) cudaFree(stack);
          set_err(fail(GDF_ERR_SHAPE, "%s: the architecture addresses row %d of a %lld-row weight", key.c_str(), v,
                       (long long)R));
          return out;
        }
    int* idx_dev = nullptr;
    if (idx) {
      cudaMalloc(&idx_dev, idx->size() * 4);
      cudaMemcpy(idx_dev, idx->data(), idx->size() * 4, cudaMemcpyHostToDevice);
    }
    bf16* dst = static_cast<bf16*>(dev_alloc((size_t)R_out * K * 2));
    float* u = static_cast<float*>(dev_alloc((size_t)R_out * 4));
    float* c = static_cast<float*>(dev_alloc((size_t)R_out * 4));
    if (dst && u && c) {
      pack_rows_scaled_kernel<<<1024, 256>>>(src, K, idx_dev, R_out, gamma, dst);
      ln_fold_consts_kernel<<<(R_out + 7) / 8, 256>>>(src, dst, K, idx_dev, R_out, beta, bias, u, c);
      cudaDeviceSynchronize();
    }
    if (stack) cudaFree(stack);
    if (idx_dev) cudaFree(idx_dev);
    h->packed[key] = dst;
    h->packed[key + "#u"] = u;
    h->packed[key + "#c"] = c;
    note_mat(dst, R_out, K);
    note_vec(u, R_out);
    note_vec(c, R_out);
    out.w = dst;
    out.u = u;
    out.c = c;
    return out;
  }
  const bf16* lin(const std::string& name) { return rows_bf16(name + "#bf16", {name}, nullptr); }
  // fp32 vectors concatenated (biases of row-concatenated projections)
  const float* f32_cat(const std::string& key, const std::vector<std::string>& names) {
    auto it = h->packed.find(key);
    if (it != h->packed.end()) return static_cast<const float*>(it->second);
    int64_t total = 0;
    for (auto& n : names) {
      const RawW* r = raw(n);
      if (!r || !src_ptr(r, n)) return nullptr;
      total += r->numel;
    }
    float* d = static_cast<float*>(dev_alloc((size_t)total * 4));
    note_vec(d, total);
    if (d) {
      int64_t off = 0;
      for (auto& n : names) {
        const RawW* r = raw(n);
        cudaMemcpy(d + off, r->ptr, (size_t)r->numel * 4, cudaMemcpyDeviceToDevice);
        off += r->numel;
      }
    }
    h->packed[key] = d;
    return d;
  }
  const float* f32_gather(const std::string& key, const std::string& name, const std::vector<int>& idx) {
    auto it = h->packed.find(key);
    if (it != h->packed.end()) return static_cast<const float*>(it->second);
    const RawW* r = raw(name);
    if (!r || !src_ptr(r, name)) return nullptr;
    for (int v : idx)
      if (v >= r->numel) {
        set_err(fail(GDF_ERR_SHAPE, "%s: the architecture addresses element %d of %lld", name.c_str(), v,
                     (long long)r->numel));
        return nullptr;
      }
    int* idx_dev = nullptr;
    cudaMalloc(&idx_dev, idx.size() * 4);
    cudaMemcpy(idx_dev, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
    float* dst = static_cast<float*>(dev_alloc(idx.size() * 4));
    if (dst) {
      gather_f32_kernel<<<((int)idx.size() + 255) / 256, 256>>>(r->ptr, idx_dev, (int)idx.size(), dst);
      cudaDeviceSynchronize();
    }
    cudaFree(idx_dev);
    h->packed[key] = dst;
    note_vec(dst, (int64_t)idx.size());
    return dst;
  }
  // conv weight OIHW -> bf16 [O_pad][k_pad]
  const bf16* conv_w(const std::string& name, int* o_pad_out, int k_pad = 0) {
    const RawW* r = raw(name);
    if (!r) return nullptr;
    if (r->shape.size() != 4) {
      set_err(fail(GDF_ERR_SHAPE, "%s is not a conv weight", name.c_str()));
      return nullptr;
    }
    const int O = (int)r->shape[0], I = (int)r->shape[1], kh = (int)r->shape[2], kw = (int)r->shape[3];
    const int O_pad = (O + 15) / 16 * 16;
    if (k_pad == 0) k_pad = kh * kw * I;
    if (o_pad_out) *o_pad_out = O_pad;
    const std::string key = name + "#conv" + std::to_string(k_pad);
    auto it = h->packed.find(key);
    if (it != h->packed.end()) return static_cast<const bf16*>(it->second);
    if (!src_ptr(r, name)) return nullptr;
    if (k_pad < kh * kw * I) {
      set_err(fail(GDF_ERR_SHAPE, "%s: %d x %d x %d taps do not fit the K = %d the architecture asks for", name.c_str(),
                   kh, kw, I, k_pad));
      return nullptr;
    }
    bf16* dst = static_cast<bf16*>(dev_alloc((size_t)O_pad * k_pad * 2));
    if (dst) {
      launch_pack_conv_weight(r->ptr, dst, O, O_pad, I, kh, kw, k_pad, 0);
      cudaDeviceSynchronize();
    }
    h->packed[key] = dst;
    note_mat(dst, O_pad, k_pad);
    return dst;
  }

  // ---- buffers -------------------------------------------------------------------------------
  bf16* buf(long long rows, int cols) {
    if (dry) return nullptr;
    void* p = h->pool.acquire((size_t)rows * cols * 2);
    if (!p) set_err(fail(GDF_ERR_CUDA, "out of device memory for a %lld x %d activation", rows, cols));
    return static_cast<bf16*>(p);
  }
  float* fbuf(long long n) {
    if (dry) return nullptr;
    void* p = h->pool.acquire((size_t)n * 4);
    if (!p) set_err(fail(GDF_ERR_CUDA, "out of device memory"));
    return static_cast<float*>(p);
  }
  void rel(void* p) {
    if (!dry) h->pool.release(p);
  }

  // ---- capture slots ------------------------------------------------------------------------
  // Registers a capture site (execution order) and returns its arena offset, or -1 when not requested.
  int64_t site(const std::string& id, int C, int H, int W) {
    if (dry) return -1;
    h->sites.push_back({id, C, H, W});
    auto it = h->requested.find(id);
    if (it == h->requested.end()) return -1;
    gdf_slot& s = h->slots[it->second];
    if (s.offset_bytes >= 0) return s.offset_bytes;  // duplicate id in the caller's list
    s.offset_bytes = h->arena_bytes;
    s.channels = C;
    s.height = H;
    s.width = W;
    s.order = (int)h->sites.size() - 1;
    h->arena_bytes += (((int64_t)h->B * H * W * C * 2) + 255) & ~int64_t(255);
    return s.offset_bytes;
  }

  // ---- op emission --------------------------------------------------------------------------
  // capture destinations live in the caller's arena, whose base is only known at run time: the plan records
  // offsets, the Epilogue carries placeholder pointers, and the launch closure patches pointers + TMA maps
  static void apply_caps(Epilogue& e, const Caps& caps, int n_out) {
    __half* const placeholder = reinterpret_cast<__half*>(uintptr_t(256));
    e.defer_capture_maps = true;
    if (caps.pre >= 0) {
      e.cap_pre = placeholder;
      e.ld_cap_pre = n_out;
    }
    e.num_cap = caps.n;
    for (int i = 0; i < caps.n; ++i) {
      e.cap[i].ptr = caps.post[i] >= 0 ? placeholder : nullptr;
      e.cap[i].col_begin = caps.c0[i];
      e.cap[i].col_end = caps.c1[i];
      e.cap[i].ld = caps.c1[i] - caps.c0[i];
    }
  }
  void push_gemm(GemmLaunch& g, const Caps& caps, int at = -1) {   // at >= 0: replaces the placeholder op at that index
    GemmParams& p = g.p;
    {
      char lbl[192];
      snprintf(lbl, sizeof(lbl), "gemm mode%d M=%d N=%d K=%d bn=%d cg=%d st=%d batch=%d act=%d res=%d caps=%d pre=%d tma=%d ksplit=%d tiles=%d",
               p.a_mode, p.M, p.N, p.K, p.block_n, p.cta_group, p.num_stages, p.batch, p.act, p.residual ? 1 : 0, caps.n,
               caps.pre >= 0 ? 1 : 0, p.tma_store, p.sk_pieces > 1 ? p.sk_pieces : 0, p.batch * p.num_m_tiles * p.num_n_tiles);
      ops->tag(kKindGemm, 2.0 * (double)p.M * (double)p.K * (double)(p.act == kActGeglu ? 2 * p.n_out : p.n_out) *
                              (double)p.batch, lbl);
    }
    struct Cached {   // capture maps for the two most recent arena bases (callers alternate between arenas)
      char* base[2] = {nullptr, nullptr};
      GemmLaunch g[2];
      int next = 0;
    };
    auto cache = std::make_shared<Cached>();
    const GemmLaunch gl = g;
    const Caps c = caps;
    const bool has_caps = (c.pre >= 0) || (c.n > 0);
    Op fn = [gl, c, cache, has_caps](const RunCtx& rc) -> int {
      if (!has_caps) {
        OP_CUDA(launch_gemm(gl, rc.stream));
        return 0;
      }
      int slot = -1;
      for (int i = 0; i < 2; ++i)
        if (cache->base[i] == rc.arena) slot = i;
      if (slot < 0) {
        slot = cache->next;
        cache->next ^= 1;
        GemmLaunch& t = cache->g[slot];
        t = gl;
        t.p.cap_pre = c.pre >= 0 ? reinterpret_cast<__half*>(rc.arena + c.pre) : nullptr;
        for (int i = 0; i < c.n; ++i)
          t.p.cap[i].ptr = c.post[i] >= 0 ? reinterpret_cast<__half*>(rc.arena + c.post[i]) : nullptr;
        GDF_TRY(build_capture_maps(&t));
        cache->base[slot] = rc.arena;
      }
      OP_CUDA(launch_gemm(cache->g[slot], rc.stream));
      return 0;
    };
    if (at < 0) {
      ops->push_back(std::move(fn));
    } else {
      ops->fns[at] = std::move(fn);
      ops->kinds[at] = ops->cur_kind;
      ops->flops[at] = ops->cur_flops;
      ops->labels[at] = ops->cur_label;
      ops->tag(kKindOther, 0.0);
    }
  }
  static constexpr long long kSkWsFloats = 74LL * 2 * 128 * 256;   // pieces <= resident CTA groups (74 pairs / 148 CTAs)
  static constexpr int kSkCntLen = 74 * 2 * 8;
  bool sk_workspace() {
    // off by default: measured slower than whole tiles on every SDXL shape (ops_gemm.cu plan_k_split)
    static const bool on = [] { const char* e = getenv("GDF_STREAM_K"); return e && e[0] == '1'; }();
    if (!on) return false;
    if (h->sk_ws && h->sk_cnt) return true;
    if (cudaMalloc(&h->sk_ws, kSkWsFloats * 4) != cudaSuccess || cudaMalloc(&h->sk_cnt, kSkCntLen * 4) != cudaSuccess ||
        cudaMemset(h->sk_cnt, 0, kSkCntLen * 4) != cudaSuccess) {
      cudaGetLastError();
      if (h->sk_ws) cudaFree(h->sk_ws);
      if (h->sk_cnt) cudaFree(h->sk_cnt);
      h->sk_ws = nullptr;
      h->sk_cnt = nullptr;
      return false;
    }
    return true;
  }
  void linear(const bf16* A, long long M, int K, int lda, const bf16* W, int N, const Epilogue& e0,
              const Caps& caps = Caps(), int batch = 1, long long abs = 0, long long wbs = 0, int ldw = 0,
              int block_n = 0, int at = -1) {
    if (err) return;
    if (!check_mat(W, N, ldw ? ldw : K, "linear weight")) return;
    if (!check_vec(e0.bias, (e0.n_out > 0 && e0.act != kActGeglu) ? e0.n_out : N, "linear bias")) return;
    if (dry) return;
    Epilogue e = e0;
    apply_caps(e, caps, e.n_out > 0 ? e.n_out : (e.act == kActGeglu ? N / 2 : N));
    if (batch == 1 && M >= 4096 && sk_workspace()) {   // K-split of the last partial wave (ops_gemm.cu plan_k_split)
      e.sk_ws = h->sk_ws;
      e.sk_ws_floats = kSkWsFloats;
      e.sk_cnt = h->sk_cnt;
      e.sk_cnt_len = kSkCntLen;
    }
    GemmLaunch g;
    if (int r = build_linear(&g, A, M, K, lda, W, N, ldw ? ldw : K, e, batch, abs, wbs, block_n)) {
      set_err(r);
      return;
    }
    push_gemm(g, caps, at);
  }
  void conv3(const bf16* X, int B, int Hin, int Win, int Cin, const bf16* Wp, int Npad, int stride, int pad_lo,
             const Epilogue& e0, const Caps& caps = Caps()) {
    if (err) return;
    if (!check_mat(Wp, Npad, 9LL * Cin, "conv3x3 weight")) return;
    if (!check_vec(e0.bias, e0.n_out > 0 ? e0.n_out : Npad, "conv3x3 bias")) return;
    if (dry) return;
    Epilogue e = e0;
    apply_caps(e, caps, e.n_out > 0 ? e.n_out : Npad);
    GemmLaunch g;
    if (int r = build_conv3x3(&g, X, B, Hin, Win, Cin, Wp, Npad, stride, pad_lo, e)) {
      set_err(r);
      return;
    }
    push_gemm(g, caps);
  }
  void groupnorm(const bf16* x, bf16* y, const std::string& prefix, int B, int HW, int C, int G, float eps,
                 bool silu) {
    const float* gm = f32(prefix + ".weight");
    const float* bt = f32(prefix + ".bias");
    if (err || !check_vec(gm, C, "GroupNorm weight") || !check_vec(bt, C, "GroupNorm bias")) return;
    if (dry) return;
    float* ws = gn_ws;
    ops->tag(kKindGroupNorm, 0.0, "groupnorm HW=" + std::to_string(HW) + " C=" + std::to_string(C));
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_groupnorm(x, y, gm, bt, B, HW, C, G, eps, silu, ws, rc.stream));
      return 0;
    });
  }
  // GroupNorm whose statistics were accumulated by the producer's epilogue into `sums`
  void groupnorm_from_sums(const bf16* x, bf16* y, const std::string& prefix, int B, int HW, int C, int G, float eps,
                           bool silu, const float* sums) {
    const float* gm = f32(prefix + ".weight");
    const float* bt = f32(prefix + ".bias");
    if (err || !check_vec(gm, C, "GroupNorm weight") || !check_vec(bt, C, "GroupNorm bias")) return;
    if (dry) return;
    float* ws = gn_ws;
    ops->tag(kKindGroupNorm, 0.0, "groupnorm(fused stats) HW=" + std::to_string(HW) + " C=" + std::to_string(C));
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_groupnorm_from_sums(x, y, gm, bt, B, HW, C, G, eps, silu, sums, ws, rc.stream));
      return 0;
    });
  }
  void layernorm(const bf16* x, bf16* y, const std::string& prefix, long long M, int C, float eps) {
    const float* gm = f32(prefix + ".weight");
    const float* bt = f32(prefix + ".bias");
    if (err || !check_vec(gm, C, "LayerNorm weight") || !check_vec(bt, C, "LayerNorm bias")) return;
    if (dry) return;
    ops->tag(kKindLayerNorm, 0.0, "layernorm M=" + std::to_string(M) + " C=" + std::to_string(C));
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_layernorm(x, y, gm, bt, M, C, eps, nullptr, nullptr, 0, rc.stream));
      return 0;
    });
  }
  // LayerNorm without affine + AdaLN-single modulation y = LN(x) * (1 + scale[b]) + shift[b] (attention.py:498-503)
  void layernorm_mod(const bf16* x, bf16* y, long long M, int C, float eps, const float* scale, const float* shift,
                     int rows_per_batch) {
    if (dry || err) return;
    ops->tag(kKindLayerNorm, 0.0, "layernorm-mod M=" + std::to_string(M) + " C=" + std::to_string(C));
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_layernorm(x, y, nullptr, nullptr, M, C, eps, scale, shift, rows_per_batch, rc.stream));
      return 0;
    });
  }
  // generic-head-dim attention with the handle's optional key bias (PixArt masked cross-attention)
  void attention_bias(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* o, int ldo, int B,
                      int heads, int Nq, int Nk, float scale, int head_dim, bool use_key_bias) {
    if (dry || err) return;
    gdf_handle_s* hh = h;
    ops->tag(kKindAttention, 4.0 * B * heads * (double)Nq * (double)Nk * head_dim,
             "attention heads=" + std::to_string(heads) + " d=" + std::to_string(head_dim) + " Nq=" +
                 std::to_string(Nq) + " Nk=" + std::to_string(Nk));
    ops->push_back([=](const RunCtx& rc) -> int {
      const float* kb = (use_key_bias && hh->has_key_bias) ? hh->key_bias : nullptr;
      OP_CUDA(launch_attention_generic(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, head_dim, scale, rc.stream, kb));
      return 0;
    });
  }
  void attention(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* o, int ldo, int B,
                 int heads, int Nq, int Nk, float scale, int v_f16, int head_dim) {
    if (dry || err) return;
    if (head_dim != 64) {
      ops->tag(kKindAttention, 4.0 * B * heads * (double)Nq * (double)Nk * head_dim,
               "attention heads=" + std::to_string(heads) + " d=" + std::to_string(head_dim) + " Nq=" +
                   std::to_string(Nq) + " Nk=" + std::to_string(Nk));
      ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_attention_generic(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, head_dim, scale, rc.stream));
        return 0;
      });
      return;
    }
    ops->tag(kKindAttention, 4.0 * B * heads * (double)Nq * (double)Nk * 64.0,
             "attention heads=" + std::to_string(heads) + " Nq=" + std::to_string(Nq) + " Nk=" + std::to_string(Nk));
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_attention64(q, ldq, k, ldk, v, ldv, o, ldo, B, heads, Nq, Nk, scale, v_f16, rc.stream));
      return 0;
    });
  }
  // Is this id (a feature id or an internal "#attnmean:<block>-<self|cross>" id) part of the current plan?
  bool wants(const std::string& id) const { return !dry && h->requested.count(id) != 0; }
  // Slow path with materialised probabilities (AttnStoreProcessor, feature/components/attention.py:165-263): the
  // (B, heads, Nq, Nk) map goes to the `<block>-<kind>-map` slot (or a scratch buffer), its head mean to the internal
  // `#attnmean:<block>-<kind>` slot that the host aggregates into the `attn` feature (diffusion_feature.py:488-500).
  void attention_probs(const std::string& block_id, const char* kind, const bf16* q, int ldq, const bf16* k, int ldk,
                       const bf16* v, int ldv, int v_f16, bf16* o, int ldo, int B, int heads, int Nq, int Nk, float scale,
                       int head_dim, bool use_key_bias = false) {
    if (dry || err) return;
    gdf_handle_s* hh = h;
    const int64_t map_off = site(block_id + "-" + kind + "-map", heads, Nq, Nk);
    const int64_t mean_off = site("#attnmean:" + block_id + "-" + kind, 1, Nq, Nk);
    __half* scratch = map_off < 0 ? reinterpret_cast<__half*>(buf((long long)B * heads * Nq, Nk)) : nullptr;
    if (err) return;
    // Tensor-core form (self maps: Nk a multiple of 8, no key bias): per image, S = scale Q K^T for every head as ONE
    // batched tcgen05 GEMM (batch = heads, batch stride = head_dim columns) into fp32 scores, softmax -> fp16
    // probabilities straight into the map's arena slot, V^T in fp16, O = P V as a second batched GEMM on the fp16
    // operand path. The CUDA-core kernel below stays for the text cross maps (77 keys: rows of 154 bytes are not TMA
    // addressable, and they are 50x smaller) and for PixArt's key bias. GDF_MAPS_TC=0 forces it everywhere (A/B).
    static const bool maps_tc = [] { const char* e = getenv("GDF_MAPS_TC"); return !(e && e[0] == '0'); }();
    if (maps_tc && !use_key_bias && !v_f16 && Nk % 8 == 0 && Nk >= 64 && Nk <= 8192 && head_dim % 8 == 0 && ldq % 8 == 0 &&
        ldk % 8 == 0 && ldo % 8 == 0) {
      float* S = fbuf((long long)heads * Nq * Nk);
      __half* vt = reinterpret_cast<__half*>(buf((long long)heads * head_dim, Nk));
      if (err) return;
      struct Cached { __half* P = nullptr; std::vector<GemmLaunch> g; };
      auto cache = std::make_shared<Cached>();
      ops->tag(kKindAttention, 4.0 * B * heads * (double)Nq * (double)Nk * head_dim,
               "attention-probs(tcgen05 GEMMs) heads=" + std::to_string(heads) + " d=" + std::to_string(head_dim) + " Nq=" +
                   std::to_string(Nq) + " Nk=" + std::to_string(Nk));
      ops->push_back([=](const RunCtx& rc) -> int {
        __half* P = map_off >= 0 ? reinterpret_cast<__half*>(rc.arena + map_off) : scratch;
        if (cache->P != P) {   // launches hold the arena address: built once per arena
          cache->g.assign(2 * (size_t)B, GemmLaunch());
          for (int b = 0; b < B; ++b) {
            Epilogue es;
            es.alpha = scale;
            es.out_f32 = S;
            es.ld_out_f32 = Nk;
            es.out_batch_stride = (long long)Nq * Nk;
            GDF_TRY(build_linear(&cache->g[2 * b], q + (long long)b * Nq * ldq, Nq, head_dim, ldq, k + (long long)b * Nk * ldk,
                                 Nk, ldk, es, heads, head_dim, head_dim));
            Epilogue eo;
            eo.in_f16 = true;
            eo.out = o + (long long)b * Nq * ldo;
            eo.ld_out = ldo;
            eo.out_batch_stride = head_dim;
            GDF_TRY(build_linear(&cache->g[2 * b + 1],
                                 reinterpret_cast<const bf16*>(P + (long long)b * heads * Nq * Nk), Nq, Nk, Nk,
                                 reinterpret_cast<const bf16*>(vt), head_dim, Nk, eo, heads, (long long)Nq * Nk,
                                 (long long)head_dim * Nk));
          }
          cache->P = P;
        }
        for (int b = 0; b < B; ++b) {
          OP_CUDA(launch_gemm(cache->g[2 * b], rc.stream));
          OP_CUDA(launch_softmax_rows_f32_f16(S, P + (long long)b * heads * Nq * Nk, (long long)heads * Nq, Nk, rc.stream));
          OP_CUDA(launch_transpose_v_f16(v + (long long)b * Nk * ldv, ldv, vt, heads, Nk, head_dim, rc.stream));
          OP_CUDA(launch_gemm(cache->g[2 * b + 1], rc.stream));
        }
        if (mean_off >= 0)
          OP_CUDA(launch_head_mean(P, reinterpret_cast<__half*>(rc.arena + mean_off), B, heads, (long long)Nq * Nk,
                                   rc.stream));
        return 0;
      });
      rel(S);
      rel(vt);
      if (scratch) rel(scratch);
      return;
    }
    ops->tag(kKindAttention, 4.0 * B * heads * (double)Nq * (double)Nk * head_dim,
             "attention-probs heads=" + std::to_string(heads) + " d=" + std::to_string(head_dim) + " Nq=" +
                 std::to_string(Nq) + " Nk=" + std::to_string(Nk));
    ops->push_back([=](const RunCtx& rc) -> int {
      __half* P = map_off >= 0 ? reinterpret_cast<__half*>(rc.arena + map_off) : scratch;
      const float* kb = (use_key_bias && hh->has_key_bias) ? hh->key_bias : nullptr;   // PixArt caption mask
      OP_CUDA(launch_attention_probs(q, ldq, k, ldk, v, ldv, v_f16, o, ldo, P, B, heads, Nq, Nk, head_dim, scale,
                                     rc.stream, kb));
      if (mean_off >= 0)
        OP_CUDA(launch_head_mean(P, reinterpret_cast<__half*>(rc.arena + mean_off), B, heads, (long long)Nq * Nk,
                                 rc.stream));
      return 0;
    });
    if (scratch) rel(scratch);
  }
  // Time-embedding projections of the resnets (Linear(SiLU(temb)), resnet.py:354-355): all of them read the same emb and
  // none depends on an activation, so they run as ONE grouped GEMV where the first resnet would have launched its own
  // (0.57 ms of 21 cold fp32 GEMVs inside the dependency chain of the SDXL step -> one launch). The table is uploaded
  // when the op list is complete (finish_temb_group); GDF_TEMB_GROUP=0 keeps one launch per resnet.
  struct TembGroup {
    std::vector<GroupedLinearItem> items;
    const float* x = nullptr;
    int B = 0, K = 0, rows = 0;
    GroupedLinearItem* dev = nullptr;
  };
  std::shared_ptr<TembGroup> temb_group;
  // All outputs of the group are written at the head of the forward, long before most resnets run, so they cannot be
  // ordinary pool buffers (the pool recycles a buffer as soon as its last consumer has been EMITTED: a later resnet's
  // output buffer could alias an activation that is still live when the grouped launch writes it). One slab, acquired
  // before the first layer and never released, is carved in emission order.
  float* temb_slab = nullptr;
  long long temb_slab_floats = 0, temb_slab_used = 0;
  void temb_reserve(long long floats) {
    temb_slab = nullptr;
    temb_slab_floats = temb_slab_used = 0;
    if (dry || floats <= 0) return;
    temb_slab = fbuf(floats);
    temb_slab_floats = temb_slab ? floats : 0;
  }
  float* temb_out(long long floats) {
    if (!temb_slab || temb_slab_used + floats > temb_slab_floats) return nullptr;
    float* p = temb_slab + temb_slab_used;
    temb_slab_used += floats;
    return p;
  }
  bool temb_projection(const float* x, const std::string& prefix, float* y, int B, int K, int N) {
    static const bool on = [] { const char* e = getenv("GDF_TEMB_GROUP"); return !(e && e[0] == '0'); }();
    if (!on || dry || (size_t)B * K * 4 > 48 * 1024) return false;
    const float* W = f32(prefix + ".weight");
    const float* bi = f32(prefix + ".bias");
    if (err || !check_vec(W, (int64_t)N * K, "time-embedding projection weight") || !check_vec(bi, N, "time-embedding projection bias"))
      return true;
    if (!temb_group) {
      temb_group = std::make_shared<TembGroup>();
      temb_group->x = x;
      temb_group->B = B;
      temb_group->K = K;
      std::shared_ptr<TembGroup> grp = temb_group;
      ops->tag(kKindOther, 0.0, "time-embedding projections of every resnet (grouped GEMV)");
      ops->push_back([grp](const RunCtx& rc) -> int {
        OP_CUDA(launch_grouped_small_linear(grp->dev, (int)grp->items.size(), grp->x, grp->B, grp->K, grp->rows, 1,
                                            rc.stream));
        return 0;
      });
    }
    if (temb_group->x != x || temb_group->K != K || temb_group->B != B) return false;   // a different embedding: own launch
    GroupedLinearItem it;
    it.W = W;
    it.bias = bi;
    it.y = y;
    it.N = N;
    it.row0 = temb_group->rows;
    temb_group->rows += N;
    temb_group->items.push_back(it);
    return true;
  }
  void finish_temb_group() {
    if (!temb_group || dry || err) return;
    const size_t bytes = temb_group->items.size() * sizeof(GroupedLinearItem);
    temb_group->dev = static_cast<GroupedLinearItem*>(dev_alloc(bytes));   // weight-lifetime storage, NOT the activation pool
    if (temb_group->dev) cudaMemcpy(temb_group->dev, temb_group->items.data(), bytes, cudaMemcpyHostToDevice);
    temb_group.reset();
  }
  void small_linear(const float* x, const std::string& prefix, float* y, int B, int K, int N, bool silu_in,
                    bool silu_out) {
    const float* W = f32(prefix + ".weight");
    const float* b = f32(prefix + ".bias");
    if (err || !check_vec(W, (int64_t)N * K, "conditioning MLP weight") || !check_vec(b, N, "conditioning MLP bias"))
      return;
    if (dry) return;
    ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_small_linear(x, W, b, y, B, K, N, silu_in, silu_out, rc.stream));
      return 0;
    });
  }
};

// ----------------------------------------------------------------------------------------- layers
struct Dest {      // where a layer's output goes (gn_sums: statistics requested for the GroupNorm that consumes it)
  bf16* out = nullptr; int ld = 0;    // primary (may be a slice of a skip-concat buffer)
  bf16* out2 = nullptr; int ld2 = 0;  // optional second copy (skip-concat slice on the down path)
  float* gn_sums = nullptr;
};

// ResnetBlock2D (resnet.py:320-379). x: [B*H*W, Cin] contiguous. temb: fp32 [B, temb_ch] or null (VAE).
static void emit_resnet(Builder& b, const std::string& wp, const std::string& fid, const bf16* x, int B, int H, int W,
                        int Cin, int Cout, const float* emb, int temb_ch, int groups, float eps, const Dest& d,
                        const float* x_sums = nullptr) {
  if (b.ops) b.ops->scope = fid.empty() ? wp : fid;   // NVTX range of this block (GDF_NVTX=1)
  // x_sums: GroupNorm statistics of x left by its producer's epilogue (null: norm1 runs its own statistics pass)
  const long long M = (long long)B * H * W;
  bf16* t1 = b.buf(M, Cin);
  if (x_sums) b.groupnorm_from_sums(x, t1, wp + ".norm1", B, H * W, Cin, groups, eps, true, x_sums);
  else b.groupnorm(x, t1, wp + ".norm1", B, H * W, Cin, groups, eps, true);
  float* tproj = nullptr;
  bool tproj_grouped = false;   // written at the head of the op list: the buffer must not be recycled before this resnet
  if (temb_ch > 0) {   // (emb itself is null during the dry weight walk)
    tproj = b.temb_out((long long)B * Cout);
    if (tproj && b.temb_projection(emb, wp + ".time_emb_proj", tproj, B, temb_ch, Cout)) {
      tproj_grouped = true;
    } else {
      tproj = b.fbuf((long long)B * Cout);
      b.small_linear(emb, wp + ".time_emb_proj", tproj, B, temb_ch, Cout, true, false);  // Linear(SiLU(temb))
    }
  }
  int npad = 0;
  const bf16* w1 = b.conv_w(wp + ".conv1.weight", &npad);
  bf16* t2 = b.buf(M, Cout);
  float* sums2 = (npad == Cout && b.gn_fusable(Cout, groups, (long long)H * W)) ? b.gn_slot(B, groups) : nullptr;
  {
    Epilogue e;
    e.bias = b.f32_pad(wp + ".conv1.bias", npad);
    e.row_batch_bias = tproj;
    e.rows_per_batch = H * W;
    e.n_out = Cout;
    e.out = t2;
    e.ld_out = Cout;
    Builder::want_gn_stats(e, sums2, Cout, groups, (long long)H * W);
    b.conv3(t1, B, H, W, Cin, w1, npad, 1, 1, e);
  }
  b.rel(t1);
  bf16* t3 = b.buf(M, Cout);
  if (sums2) b.groupnorm_from_sums(t2, t3, wp + ".norm2", B, H * W, Cout, groups, eps, true, sums2);
  else b.groupnorm(t2, t3, wp + ".norm2", B, H * W, Cout, groups, eps, true);
  b.rel(t2);
  const bf16* res = x;
  bf16* sc = nullptr;
  if (Cin != Cout) {  // 1x1 conv_shortcut on the raw input
    int np2 = 0;
    const bf16* ws = b.conv_w(wp + ".conv_shortcut.weight", &np2);
    sc = b.buf(M, Cout);
    Epilogue e;
    e.bias = b.f32_pad(wp + ".conv_shortcut.bias", np2);
    e.n_out = Cout;
    e.out = sc;
    e.ld_out = Cout;
    b.linear(x, M, Cin, Cin, ws, np2, e);
    res = sc;
  }
  const bf16* w2 = b.conv_w(wp + ".conv2.weight", &npad);
  {
    Epilogue e;
    e.bias = b.f32_pad(wp + ".conv2.bias", npad);
    e.n_out = Cout;
    e.residual = res;
    e.ld_res = Cout;
    e.out = d.out;
    e.ld_out = d.ld;
    e.out2 = d.out2;
    e.ld_out2 = d.ld2;
    Builder::want_gn_stats(e, d.gn_sums, Cout, groups, (long long)H * W);
    Caps caps;
    if (!fid.empty()) {
      caps.pre = b.site(fid + "-increment", Cout, H, W);
      const int64_t o = b.site(fid + "-out", Cout, H, W);
      caps.add(o, 0, Cout);
    }
    b.conv3(t3, B, H, W, Cout, w2, npad, 1, 1, e, caps);
  }
  b.rel(t3);
  b.rel(sc);
  if (!tproj_grouped) b.rel(tproj);   // grouped outputs live in the slab reserved at the head of the op list
}

// BasicTransformerBlock (attention.py:469-592) on hs [M, C]; returns the new hidden-state buffer.
// sums_in: row statistics of hs (written by its producer), sums_out: where the block's last GEMM adds the statistics of
// its output for the next block's norm1 (null: not needed). Both null when the LayerNorms are not folded.
static bf16* emit_tblock(Builder& b, const std::string& wp, const std::string& fid, bf16* hs, int B, int N, int C,
                         int heads, int ctx_dim, int hw, const float* sums_in = nullptr, float* sums_out = nullptr) {
  if (b.ops) b.ops->scope = fid.empty() ? wp : fid;
  gdf_handle_s* h = b.h;
  const long long M = (long long)B * N;
  const float scale = 1.f / sqrtf((float)(C / heads));
  // ---- self attention
  const int hd = C / heads;
  const bool probs_self = b.wants(fid + "-self-map") || b.wants("#attnmean:" + fid + "-self");
  const bool probs_cross = b.wants(fid + "-cross-map") || b.wants("#attnmean:" + fid + "-cross");
  const bool self_tc = !probs_self && (hd == 64) && attention_uses_tcgen05(N);
  const bool fold = b.ln_fold;
  float* sums1 = fold ? b.ln_rows(M) : nullptr;   // statistics of hs1 / hs2 (inputs of norm2 / norm3)
  float* sums2 = fold ? b.ln_rows(M) : nullptr;
  const std::vector<std::string> qkv_names = {wp + ".attn1.to_q.weight", wp + ".attn1.to_k.weight",
                                              wp + ".attn1.to_v.weight"};
  bf16* n1 = nullptr;
  const bf16* wqkv = nullptr;
  Builder::LnW lw1;
  if (fold) {
    lw1 = b.rows_bf16_ln(wp + ".attn1#qkv_ln", qkv_names, nullptr, wp + ".norm1", "");
    wqkv = lw1.w;
  } else {
    n1 = b.buf(M, C);
    b.layernorm(hs, n1, wp + ".norm1", M, C, 1e-5f);
    wqkv = b.rows_bf16(wp + ".attn1#qkv", qkv_names, nullptr);
  }
  bf16* qkv = b.buf(M, 3 * C);
  {
    Epilogue e;
    if (fold) {
      e.bias = lw1.c;
      e.ln_sums = sums_in;
      e.ln_u = lw1.u;
      e.ln_eps = 1e-5f;
    }
    e.out = qkv;
    e.ld_out = 3 * C;
    if (self_tc) e.out_f16_from = 2 * C;   // V columns in fp16 for the fp16 P.V product of the tcgen05 kernel
    Caps caps;
    caps.add(b.site(fid + "-self-q", C, hw, hw), 0, C);
    caps.add(b.site(fid + "-self-k", C, hw, hw), C, 2 * C);
    caps.add(b.site(fid + "-self-v", C, hw, hw), 2 * C, 3 * C);
    b.linear(fold ? hs : n1, M, C, C, wqkv, 3 * C, e, caps);
  }
  if (n1) b.rel(n1);
  bf16* ao = b.buf(M, C);
  if (probs_self)
    b.attention_probs(fid, "self", qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, 0, ao, C, B, heads, N, N, scale, hd);
  else
    b.attention(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, ao, C, B, heads, N, N, scale, self_tc ? 1 : 0, hd);
  b.rel(qkv);
  bf16* hs1 = b.buf(M, C);
  {
    Epilogue e;
    e.bias = b.f32(wp + ".attn1.to_out.0.bias");
    e.residual = hs;
    e.ld_res = C;
    e.out = hs1;
    e.ld_out = C;
    e.row_sums = sums1;
    b.linear(ao, M, C, C, b.lin(wp + ".attn1.to_out.0.weight"), C, e);
  }
  b.rel(ao);
  b.rel(hs);
  // ---- cross attention (context K/V: B*ctx_len rows; reference drops cross-k / cross-v, feature_extractor.py:38)
  bf16* n2 = nullptr;
  const bf16* wq2 = nullptr;
  Builder::LnW lw2;
  if (fold) {
    lw2 = b.rows_bf16_ln(wp + ".attn2#q_ln", {wp + ".attn2.to_q.weight"}, nullptr, wp + ".norm2", "");
    wq2 = lw2.w;
  } else {
    n2 = b.buf(M, C);
    b.layernorm(hs1, n2, wp + ".norm2", M, C, 1e-5f);
    wq2 = b.lin(wp + ".attn2.to_q.weight");
  }
  bf16* q2 = b.buf(M, C);
  {
    Epilogue e;
    if (fold) {
      e.bias = lw2.c;
      e.ln_sums = sums1;
      e.ln_u = lw2.u;
      e.ln_eps = 1e-5f;
    }
    e.out = q2;
    e.ld_out = C;
    Caps caps;
    caps.add(b.site(fid + "-cross-q", C, hw, hw), 0, C);
    b.linear(fold ? hs1 : n2, M, C, C, wq2, C, e, caps);
  }
  if (n2) b.rel(n2);
  const long long Mc = (long long)B * h->ctx_len;
  bf16* kv = nullptr;
  int ldkv = 2 * C;
  if (b.kv_batched) {
    b.kv_names.push_back(wp + ".attn2.to_k.weight");
    b.kv_names.push_back(wp + ".attn2.to_v.weight");
    kv = b.kv_all ? b.kv_all + b.kv_col : nullptr;
    ldkv = b.kv_ld;
    b.kv_col += 2 * C;
  } else {
    const bf16* wkv = b.rows_bf16(wp + ".attn2#kv", {wp + ".attn2.to_k.weight", wp + ".attn2.to_v.weight"}, nullptr);
    kv = b.buf(Mc, 2 * C);
    Epilogue e;
    e.out = kv;
    e.ld_out = 2 * C;
    b.linear(h->ctx_bf16, Mc, ctx_dim, ctx_dim, wkv, 2 * C, e);
  }
  bf16* ao2 = b.buf(M, C);
  if (probs_cross)
    b.attention_probs(fid, "cross", q2, C, kv, ldkv, kv + C, ldkv, 0, ao2, C, B, heads, N, h->ctx_len, scale, hd);
  else
    b.attention(q2, C, kv, ldkv, kv + C, ldkv, ao2, C, B, heads, N, h->ctx_len, scale, 0, hd);
  b.rel(q2);
  if (!b.kv_batched) b.rel(kv);
  bf16* hs2 = b.buf(M, C);
  {
    Epilogue e;
    e.bias = b.f32(wp + ".attn2.to_out.0.bias");
    e.residual = hs1;
    e.ld_res = C;
    e.out = hs2;
    e.ld_out = C;
    e.row_sums = sums2;
    b.linear(ao2, M, C, C, b.lin(wp + ".attn2.to_out.0.weight"), C, e);
  }
  b.rel(ao2);
  b.rel(hs1);
  // ---- feed-forward (GEGLU), attention.py:1249-1258
  bf16* n3 = nullptr;
  if (!fold) {
    n3 = b.buf(M, C);
    b.layernorm(hs2, n3, wp + ".norm3", M, C, 1e-5f);
  }
  const int inner = 4 * C;
  const int bn = 256, half = bn / 2;
  std::vector<int> idx;
  idx.reserve(2 * inner);
  for (int t = 0; t < inner / half; ++t) {
    for (int j = 0; j < half; ++j) idx.push_back(t * half + j);
    for (int j = 0; j < half; ++j) idx.push_back(inner + t * half + j);
  }
  const bf16* w1 = nullptr;
  const float* b1 = nullptr;
  Builder::LnW lw3;
  if (fold) {
    lw3 = b.rows_bf16_ln(wp + ".ff#geglu_ln", {wp + ".ff.net.0.proj.weight"}, &idx, wp + ".norm3",
                         wp + ".ff.net.0.proj.bias");
    w1 = lw3.w;
    b1 = lw3.c;
  } else {
    w1 = b.rows_bf16(wp + ".ff#geglu", {wp + ".ff.net.0.proj.weight"}, &idx);
    b1 = b.f32_gather(wp + ".ff#geglu_bias", wp + ".ff.net.0.proj.bias", idx);
  }
  bf16* ffi = b.buf(M, inner);
  {
    Epilogue e;
    if (fold) {
      e.ln_sums = sums2;
      e.ln_u = lw3.u;
      e.ln_eps = 1e-5f;
    }
    e.act = kActGeglu;
    e.bias = b1;
    e.out = ffi;
    e.ld_out = inner;
    Caps caps;
    caps.add(b.site(fid + "-ffn-inner", inner, hw, hw), 0, inner);
    b.linear(fold ? hs2 : n3, M, C, C, w1, 2 * inner, e, caps, 1, 0, 0, 0, bn);
  }
  if (n3) b.rel(n3);
  bf16* hs3 = b.buf(M, C);
  {
    Epilogue e;
    e.bias = b.f32(wp + ".ff.net.2.bias");
    e.residual = hs2;
    e.ld_res = C;
    e.out = hs3;
    e.ld_out = C;
    e.row_sums = sums_out;
    Caps caps;
    caps.add(b.site(fid + "-out", C, hw, hw), 0, C);
    b.linear(ffi, M, inner, inner, b.lin(wp + ".ff.net.2.weight"), C, e, caps);
  }
  b.rel(ffi);
  b.rel(hs2);
  return hs3;
}

// Transformer2DModel (transformers/transformer_2d.py:403-530). x [B*hw*hw, C] contiguous (NHWC == token-major).
static void emit_vit(Builder& b, const std::string& wp, const std::string& fid, const bf16* x, int B, int hw, int C,
                     int heads, int depth, int ctx_dim, int groups, const Dest& d) {
  if (b.ops) b.ops->scope = fid.empty() ? wp : fid;
  const int N = hw * hw;
  const long long M = (long long)B * N;
  bf16* t = b.buf(M, C);
  b.groupnorm(x, t, wp + ".norm", B, N, C, groups, 1e-6f, false);
  bf16* hs = b.buf(M, C);
  // folded LayerNorms: row statistics of the current block input, added by the GEMM that produces it
  float* sums = b.ln_fold ? b.ln_rows(M) : nullptr;
  {
    // Linear (use_linear_projection) or 1x1 conv: the same [C, C] contraction in NHWC
    Epilogue e;
    e.bias = b.f32(wp + ".proj_in.bias");
    e.out = hs;
    e.ld_out = C;
    e.row_sums = sums;
    b.linear(t, M, C, C, b.lin(wp + ".proj_in.weight"), C, e);
  }
  b.rel(t);
  for (int k = 0; k < depth; ++k) {
    float* sums_next = (b.ln_fold && k + 1 < depth) ? b.ln_rows(M) : nullptr;
    hs = emit_tblock(b, wp + ".transformer_blocks." + std::to_string(k), fid + "-block" + std::to_string(k), hs, B, N,
                     C, heads, ctx_dim, hw, sums, sums_next);
    sums = sums_next;
  }
  if (b.ops) b.ops->scope = fid;
  {
    Epilogue e;
    e.bias = b.f32(wp + ".proj_out.bias");
    e.residual = x;
    e.ld_res = C;
    e.out = d.out;
    e.ld_out = d.ld;
    e.out2 = d.out2;
    e.ld_out2 = d.ld2;
    Caps caps;
    caps.add(b.site(fid + "-out", C, hw, hw), 0, C);
    b.linear(hs, M, C, C, b.lin(wp + ".proj_out.weight"), C, e, caps);
  }
  b.rel(hs);
}

// ----------------------------------------------------------------------------------------- UNet
struct Skip {
  int C, hw;
  bf16* cbuf;   // consumer's concat buffer [M, ld]
  int ld, col;  // this skip lives in columns [col, col + C)
};

static int build_unet(Builder& b) {
  gdf_handle_s* h = b.h;
  const gdf_unet_arch& a = h->ua;
  const int B = h->B, nl = a.num_levels, lpb = a.layers_per_block, G = a.norm_num_groups;
  const float eps = a.norm_eps;
  const int temb_ch = a.block_out_channels[0] * 4;
  const std::string U = "unet.";
  b.ops = &h->unet_ops;
  b.ops->scope = "denoiser-head";
  b.gn_fuse = false;   // fused GroupNorm statistics: VAE only (UNet groups of 10 / 20 / 40 channels are not supported)
  b.gn_cap = 0;

  // ---- skip bookkeeping: channels/resolution of every skip in push order, consumers in pop order
  std::vector<Skip> skips;
  {
    int hw = h->L;
    skips.push_back({a.block_out_channels[0], hw, nullptr, 0, 0});
    for (int i = 0; i < nl; ++i) {
      for (int j = 0; j < lpb; ++j) skips.push_back({a.block_out_channels[i], hw, nullptr, 0, 0});
      if (i != nl - 1) {
        hw /= 2;
        skips.push_back({a.block_out_channels[i], hw, nullptr, 0, 0});
      }
    }
  }
  // up-path geometry: resnet (i, j) consumes skip index sidx with previous-h channels cprev
  struct UpRes { int cprev, cskip, cout, hw, sidx; bf16* cbuf; };
  std::vector<UpRes> upres;
  {
    int sidx = (int)skips.size() - 1;
    int prev = a.block_out_channels[nl - 1];
    for (int i = 0; i < nl; ++i) {
      const int cout = a.block_out_channels[nl - 1 - i];
      for (int j = 0; j < lpb + 1; ++j) {
        UpRes u;
        u.cprev = (j == 0) ? prev : cout;
        u.cskip = skips[sidx].C;
        u.cout = cout;
        u.hw = skips[sidx].hw;
        u.sidx = sidx--;
        u.cbuf = b.buf((long long)B * u.hw * u.hw, u.cprev + u.cskip);
        skips[u.sidx].cbuf = u.cbuf;
        skips[u.sidx].ld = u.cprev + u.cskip;
        skips[u.sidx].col = u.cprev;
        upres.push_back(u);
      }
      prev = cout;
    }
  }

  // ---- conditioning embeddings (unet_2d_condition.py:1141-1162, 910-1002), fp32
  float* temb_sin = b.fbuf((long long)B * a.block_out_channels[0]);
  float* emb1 = b.fbuf((long long)B * temb_ch);
  float* emb = b.fbuf((long long)B * temb_ch);
  if (!b.dry) {
    float* t_dev = h->t_dev;
    const int dim0 = a.block_out_channels[0];
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_timestep_embedding(t_dev, temb_sin, B, dim0, rc.stream));
      return 0;
    });
  }
  b.small_linear(temb_sin, U + "time_embedding.linear_1", emb1, B, a.block_out_channels[0], temb_ch, false, true);
  b.small_linear(emb1, U + "time_embedding.linear_2", emb, B, temb_ch, temb_ch, false, false);
  if (a.addition_time_embed_dim > 0) {
    const int td = a.addition_time_embed_dim, in_dim = a.projection_class_embeddings_input_dim;
    const int pooled_dim = in_dim - 6 * td;
    float* te = b.fbuf((long long)B * 6 * td);
    float* aug1 = b.fbuf((long long)B * temb_ch);
    float* aug = b.fbuf((long long)B * temb_ch);
    if (!b.dry) {
      gdf_handle_s* hh = h;
      float* add_in = h->add_in;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        if (!hh->pooled || !hh->time_ids_dev)
          return fail(GDF_ERR_INVALID, "this UNet needs pooled text embeds and add_time_ids (SDXL text_time)");
        OP_CUDA(launch_timestep_embedding(hh->time_ids_dev, te, B * 6, td, rc.stream));
        // add_embeds = cat([text_embeds, time_embeds], -1)  (unet_2d_condition.py:981)
        OP_CUDA(cudaMemcpy2DAsync(add_in, (size_t)in_dim * 4, hh->pooled, (size_t)pooled_dim * 4,
                                  (size_t)pooled_dim * 4, B, cudaMemcpyDeviceToDevice, rc.stream));
        OP_CUDA(cudaMemcpy2DAsync(add_in + pooled_dim, (size_t)in_dim * 4, te, (size_t)6 * td * 4, (size_t)6 * td * 4,
                                  B, cudaMemcpyDeviceToDevice, rc.stream));
        return 0;
      });
    }
    b.small_linear(h->add_in, U + "add_embedding.linear_1", aug1, B, in_dim, temb_ch, false, true);
    b.small_linear(aug1, U + "add_embedding.linear_2", aug, B, temb_ch, temb_ch, false, false);
    if (!b.dry) {
      const int n = B * temb_ch;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        add_f32_kernel<<<(n + 255) / 256, 256, 0, rc.stream>>>(emb, aug, n);
        OP_CUDA(cudaGetLastError());
        return 0;
      });
    }
  }

  {   // output slab of the grouped time-embedding projections: B x (sum of the output widths of every resnet)
    long long tot = 0;
    for (int i = 0; i < nl; ++i) tot += (long long)lpb * a.block_out_channels[i];               // down
    tot += 2ll * a.block_out_channels[nl - 1];                                                     // mid
    for (int i = 0; i < nl; ++i) tot += (long long)(lpb + 1) * a.block_out_channels[nl - 1 - i];   // up
    b.temb_reserve((long long)B * tot);
  }

  // ---- cross-attention K/V of all transformer blocks: one placeholder op here, filled in once every block has
  // registered its to_k / to_v weights (attention_processor.py:3283-3284; the context is the same for every block)
  int kv_op = -1;
  {
    long long kv_cols = 0;
    for (int i = 0; i < nl; ++i) {
      if (a.down_has_attn[i]) kv_cols += (long long)lpb * a.transformer_depth[i] * 2 * a.block_out_channels[i];
      const int li = nl - 1 - i;
      if (a.up_has_attn[i]) kv_cols += (long long)(lpb + 1) * a.transformer_depth[li] * 2 * a.block_out_channels[li];
    }
    kv_cols += (long long)a.transformer_depth[nl - 1] * 2 * a.block_out_channels[nl - 1];   // mid block
    // folded LayerNorms: 3 statistics slices [B * hw^2][2] per transformer block (block input, hs1, hs2)
    long long ln_floats = 0;
    for (int i = 0; i < nl; ++i) {
      const long long Mi = (long long)B * (h->L >> i) * (h->L >> i);
      if (a.down_has_attn[i]) ln_floats += (long long)lpb * a.transformer_depth[i] * 3 * 2 * Mi;
      const int li = nl - 1 - i;
      const long long Mu = (long long)B * (h->L >> li) * (h->L >> li);
      if (a.up_has_attn[i]) ln_floats += (long long)(lpb + 1) * a.transformer_depth[li] * 3 * 2 * Mu;
    }
    ln_floats += (long long)a.transformer_depth[nl - 1] * 3 * 2 * B * (h->L >> (nl - 1)) * (h->L >> (nl - 1));
    {
      // Off by default: measured on B200 (profiles/r01_ln_fold_ab.md) the three consumer GEMMs of a block (K = C,
      // epilogue-latency bound) lose more (+6.3 ms / step) than the 210 LayerNorm launches cost (4.5 ms / step).
      const char* ev = getenv("GDF_LN_FOLD");
      b.ln_fold = ev && ev[0] == '1';
    }
    b.ln_off = 0;
    b.ln_total = ln_floats;
    b.ln_slab = b.ln_fold ? b.fbuf(ln_floats) : nullptr;
    if (b.ln_fold && !b.dry) {
      float* slab = b.ln_slab;
      const size_t bytes = (size_t)ln_floats * 4;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(cudaMemsetAsync(slab, 0, bytes, rc.stream));
        return 0;
      });
    }
    b.kv_batched = true;
    b.kv_names.clear();
    b.kv_col = 0;
    b.kv_ld = (int)kv_cols;
    b.kv_all = b.buf((long long)B * h->ctx_len, (int)kv_cols);
    if (!b.dry) {
      kv_op = (int)b.ops->size();
      b.ops->push_back([](const RunCtx&) -> int { return 0; });
    }
  }

  // ---- conv_in (unet_2d_condition.py:1169-1173): latent NHWC [B, L*L, 4] -> im2col(K=36->64) -> GEMM
  int sidx = 0;  // next skip to produce
  int hw = h->L;
  const long long M0 = (long long)B * hw * hw;
  h->unet_in_cap = b.site("unet-in", a.in_channels, hw, hw);
  bf16* cur = b.buf(M0, a.block_out_channels[0]);
  {
    bf16* col = b.buf(M0, 64);
    if (!b.dry) {
      bf16* lat = h->latent_nhwc;
      const int L = hw, cin = a.in_channels;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_im2col_small(nullptr, lat, col, B, L, L, cin, rc.stream));
        return 0;
      });
    }
    int npad = 0;
    const bf16* w = b.conv_w(U + "conv_in.weight", &npad, 64);
    Epilogue e;
    e.bias = b.f32_pad(U + "conv_in.bias", npad);
    e.n_out = a.block_out_channels[0];
    e.out = cur;
    e.ld_out = a.block_out_channels[0];
    e.out2 = skips[sidx].cbuf ? skips[sidx].cbuf + skips[sidx].col : nullptr;
    e.ld_out2 = skips[sidx].ld;
    Caps caps;
    caps.add(b.site("unet-after-conv-in", a.block_out_channels[0], hw, hw), 0, a.block_out_channels[0]);
    b.linear(col, M0, 64, 64, w, npad, e, caps);
    b.rel(col);
    ++sidx;
  }

  auto skip_dest = [&](bf16* fresh, int C) {
    Dest d;
    d.out = fresh;
    d.ld = C;
    d.out2 = skips[sidx].cbuf ? skips[sidx].cbuf + skips[sidx].col : nullptr;
    d.ld2 = skips[sidx].ld;
    ++sidx;
    return d;
  };

  // ---- down blocks
  int ch = a.block_out_channels[0];
  for (int i = 0; i < nl; ++i) {
    const int cout = a.block_out_channels[i];
    const bool attn = a.down_has_attn[i] != 0;
    for (int j = 0; j < lpb; ++j) {
      const std::string wp = U + "down_blocks." + std::to_string(i);
      const std::string fid = "down-level" + std::to_string(i) + "-repeat" + std::to_string(j);
      const long long M = (long long)B * hw * hw;
      bf16* r_out = b.buf(M, cout);
      Dest d;
      if (attn) { d.out = r_out; d.ld = cout; } else d = skip_dest(r_out, cout);
      emit_resnet(b, wp + ".resnets." + std::to_string(j), fid + "-res", cur, B, hw, hw, ch, cout, emb, temb_ch, G, eps,
                  d);
      b.rel(cur);
      cur = r_out;
      if (attn) {
        bf16* v_out = b.buf(M, cout);
        emit_vit(b, wp + ".attentions." + std::to_string(j), fid + "-vit", cur, B, hw, cout, a.num_heads[i],
                 a.transformer_depth[i], a.cross_attention_dim, G, skip_dest(v_out, cout));
        b.rel(cur);
        cur = v_out;
      }
      ch = cout;
    }
    if (i != nl - 1) {  // Downsample2D: conv3x3 stride 2, padding 1 (downsampling.py:147)
      const std::string wp = U + "down_blocks." + std::to_string(i) + ".downsamplers.0.conv";
      int npad = 0;
      const bf16* w = b.conv_w(wp + ".weight", &npad);
      const int ho = hw / 2;
      bf16* o = b.buf((long long)B * ho * ho, ch);
      const Dest d = skip_dest(o, ch);
      Epilogue e;
      e.bias = b.f32_pad(wp + ".bias", npad);
      e.n_out = ch;
      e.out = d.out;
      e.ld_out = d.ld;
      e.out2 = d.out2;
      e.ld_out2 = d.ld2;
      Caps caps;
      caps.add(b.site("down-level" + std::to_string(i) + "-downsampler-out", ch, ho, ho), 0, ch);
      b.conv3(cur, B, hw, hw, ch, w, npad, 2, 1, e, caps);
      b.rel(cur);
      cur = o;
      hw = ho;
    }
  }

  // ---- mid block: resnet -> vit -> resnet; the last resnet writes into the first up concat buffer
  {
    const long long M = (long long)B * hw * hw;
    const int li = nl - 1;
    bf16* r0 = b.buf(M, ch);
    Dest d0;
    d0.out = r0;
    d0.ld = ch;
    emit_resnet(b, U + "mid_block.resnets.0", "mid-repeat0-res", cur, B, hw, hw, ch, ch, emb, temb_ch, G, eps, d0);
    b.rel(cur);
    bf16* v = b.buf(M, ch);
    Dest dv;
    dv.out = v;
    dv.ld = ch;
    emit_vit(b, U + "mid_block.attentions.0", "mid-vit", r0, B, hw, ch, a.num_heads[li], a.transformer_depth[li],
             a.cross_attention_dim, G, dv);
    b.rel(r0);
    Dest d1;
    d1.out = upres[0].cbuf;
    d1.ld = upres[0].cprev + upres[0].cskip;
    emit_resnet(b, U + "mid_block.resnets.1", "mid-repeat1-res", v, B, hw, hw, ch, ch, emb, temb_ch, G, eps, d1);
    b.rel(v);
    cur = nullptr;
  }

  // ---- ControlNet residual inputs (unet_2d_condition.py:1236-1247, 1261-1275): every skip tensor and the mid-block
  // output take an additive residual when the caller supplied them (gdf_set_control_residuals). The skips live as
  // column slices of the up path's concat buffers and are complete by now; the captures of the down / mid path were
  // written by the producers' epilogues before this point, like the reference's gather sites.
  h->n_skips = (int)skips.size();
  h->skip_shapes.clear();
  for (auto& sk : skips) h->skip_shapes.push_back({sk.C, sk.hw});
  if (!b.dry) {
    gdf_handle_s* hh = h;
    const std::vector<Skip> sk = skips;
    bf16* mid_dst = upres[0].cbuf;
    const int mid_ld = upres[0].cprev + upres[0].cskip, mid_c = ch, mid_hw = hw;
    b.ops->tag(kKindOther, 0.0, "controlnet residual adds (no-op without residuals)");
    b.ops->push_back([=](const RunCtx& rc) -> int {
      if (hh->ctrl_down.empty() && !hh->ctrl_mid) return 0;
      for (size_t i = 0; i < sk.size() && i < hh->ctrl_down.size(); ++i) {
        if (!hh->ctrl_down[i] || !sk[i].cbuf) continue;
        const int HW = sk[i].hw * sk[i].hw;
        add_residual_nchw_kernel<<<1024, 256, 0, rc.stream>>>(sk[i].cbuf + sk[i].col, sk[i].ld, hh->ctrl_down[i], B, HW,
                                                             sk[i].C);
      }
      if (hh->ctrl_mid)
        add_residual_nchw_kernel<<<1024, 256, 0, rc.stream>>>(mid_dst, mid_ld, hh->ctrl_mid, B, mid_hw * mid_hw, mid_c);
      OP_CUDA(cudaGetLastError());
      return 0;
    });
  }

  // ---- up blocks
  size_t ur = 0;
  bf16* final_h = nullptr;
  for (int i = 0; i < nl; ++i) {
    const int li = nl - 1 - i;
    const int cout = a.block_out_channels[li];
    const bool attn = a.up_has_attn[i] != 0;
    const std::string wp = U + "up_blocks." + std::to_string(i);
    for (int j = 0; j < lpb + 1; ++j, ++ur) {
      const UpRes& u = upres[ur];
      const std::string fid = "up-level" + std::to_string(i) + "-repeat" + std::to_string(j);
      const long long M = (long long)B * hw * hw;
      const bool last_in_block = (j == lpb);
      const bool has_up = (i != nl - 1);
      // destination of this layer's final output
      Dest nd;
      bf16* fresh = nullptr;
      if (!last_in_block) {
        nd.out = upres[ur + 1].cbuf;
        nd.ld = upres[ur + 1].cprev + upres[ur + 1].cskip;
      } else {
        fresh = b.buf(M, cout);
        nd.out = fresh;
        nd.ld = cout;
      }
      if (attn) {
        bf16* r_out = b.buf(M, cout);
        Dest dr;
        dr.out = r_out;
        dr.ld = cout;
        emit_resnet(b, wp + ".resnets." + std::to_string(j), fid + "-res", u.cbuf, B, hw, hw, u.cprev + u.cskip, cout,
                    emb, temb_ch, G, eps, dr);
        emit_vit(b, wp + ".attentions." + std::to_string(j), fid + "-vit", r_out, B, hw, cout, a.num_heads[li],
                 a.transformer_depth[li], a.cross_attention_dim, G, nd);
        b.rel(r_out);
      } else {
        emit_resnet(b, wp + ".resnets." + std::to_string(j), fid + "-res", u.cbuf, B, hw, hw, u.cprev + u.cskip, cout,
                    emb, temb_ch, G, eps, nd);
      }
      b.rel(u.cbuf);
      if (last_in_block) {
        if (has_up) {  // Upsample2D: nearest x2 + conv3x3 (upsampling.py:176-193)
          bf16* up = b.buf(M * 4, cout);
          if (!b.dry) {
            const int hh = hw;
            bf16* src = fresh;
            b.ops->push_back([=](const RunCtx& rc) -> int {
              OP_CUDA(launch_upsample_nearest2x(src, up, B, hh, hh, cout, rc.stream));
              return 0;
            });
          }
          b.rel(fresh);
          hw *= 2;
          int npad = 0;
          const bf16* w = b.conv_w(wp + ".upsamplers.0.conv.weight", &npad);
          Epilogue e;
          e.bias = b.f32_pad(wp + ".upsamplers.0.conv.bias", npad);
          e.n_out = cout;
          e.out = upres[ur + 1].cbuf;
          e.ld_out = upres[ur + 1].cprev + upres[ur + 1].cskip;
          Caps caps;
          caps.add(b.site("up-level" + std::to_string(i) + "-upsampler-out", cout, hw, hw), 0, cout);
          b.conv3(up, B, hw, hw, cout, w, npad, 1, 1, e, caps);
          b.rel(up);
        } else {
          final_h = fresh;
        }
      }
    }
  }

  // ---- fill in the batched context projection (all blocks have registered their weights by now)
  {
    if (b.kv_col != b.kv_ld && !b.err)
      return b.set_err(fail(GDF_ERR_SHAPE, "cross-attention K/V columns: planned %d, emitted %d", b.kv_ld, b.kv_col));
    const bf16* wkv_all = b.rows_bf16("unet#cross_kv_all", b.kv_names, nullptr);
    if (!b.dry && !b.err) {
      Epilogue e;
      e.out = b.kv_all;
      e.ld_out = b.kv_ld;
      b.linear(h->ctx_bf16, (long long)B * h->ctx_len, a.cross_attention_dim, a.cross_attention_dim, wkv_all, b.kv_ld, e,
               Caps(), 1, 0, 0, 0, 0, kv_op);
    }
    b.kv_batched = false;
    if (b.ln_fold && b.ln_off != b.ln_total && !b.err)
      return b.set_err(fail(GDF_ERR_SHAPE, "LayerNorm statistics slab: planned %lld floats, used %lld", b.ln_total,
                            b.ln_off));
    b.ln_fold = false;
  }

  // ---- conv_norm_out + SiLU + conv_out (unet_2d_condition.py:1304-1310)
  {
    const int c0 = a.block_out_channels[0];
    const long long M = (long long)B * hw * hw;
    bf16* t = b.buf(M, c0);
    b.groupnorm(final_h, t, U + "conv_norm_out", B, hw * hw, c0, G, eps, true);
    b.rel(final_h);
    int npad = 0;
    const bf16* w = b.conv_w(U + "conv_out.weight", &npad);
    bf16* o = b.buf(M, a.out_channels);
    Epilogue e;
    e.bias = b.f32_pad(U + "conv_out.bias", npad);
    e.n_out = a.out_channels;
    e.out = o;
    e.ld_out = a.out_channels;
    Caps caps;
    caps.add(b.site("unet-out", a.out_channels, hw, hw), 0, a.out_channels);
    b.conv3(t, B, hw, hw, c0, w, npad, 1, 1, e, caps);
    b.rel(t);
    if (!b.dry) {
      const int HW = hw * hw, C = a.out_channels;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        if (rc.noise_pred_out) {
          nhwc_bf16_to_nchw_f32_kernel<<<256, 256, 0, rc.stream>>>(o, rc.noise_pred_out, B, HW, C);
          OP_CUDA(cudaGetLastError());
        }
        return 0;
      });
    }
  }
  b.finish_temb_group();
  return b.err;
}

// ----------------------------------------------------------------------------------------- PixArt DiT
// [diffusers PixArtTransformer2DModel.forward, un-vendored] around the reference's vendored BasicTransformerBlock
// (attention.py:469-592, norm_type 'ada_norm_single'); capture sites = feature_extractor.py:259-286.
static int build_dit(Builder& b) {
  gdf_handle_s* h = b.h;
  const gdf_dit_arch& a = h->da;
  const int B = h->B, p = a.patch_size, heads = a.num_heads, hd = a.head_dim;
  const int C = heads * hd, g = h->L / p, N = g * g, Lc = h->ctx_len;
  const long long M = (long long)B * N, Mc = (long long)B * Lc;
  const float eps = a.norm_eps;
  const float scale = 1.f / sqrtf((float)hd);
  const std::string T = "transformer.";
  b.ops = &h->unet_ops;
  b.ops->scope = "denoiser-head";
  b.gn_fuse = false;
  b.gn_cap = 0;
  h->unet_in_cap = -1;
  if (C % 64 != 0 || hd % 8 != 0 || hd > 160 || p * p * a.in_channels > 64 || a.caption_channels % 8 != 0)
    return b.set_err(fail(GDF_ERR_UNSUPPORTED, "DiT: hidden size %d / head_dim %d / patch %d unsupported", C, hd, p));

  // ---- conditioning: adaln_single(t) -> t6 [B, 6C], embedded timestep [B, C] (fp32)
  float* tsin = b.fbuf((long long)B * 256);
  float* e1 = b.fbuf((long long)B * C);
  float* emb = b.fbuf((long long)B * C);
  float* t6 = b.fbuf((long long)B * 6 * C);
  if (!b.dry) {
    float* t_dev = h->t_dev;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_timestep_embedding(t_dev, tsin, B, 256, rc.stream));
      return 0;
    });
  }
  b.small_linear(tsin, T + "adaln_single.emb.timestep_embedder.linear_1", e1, B, 256, C, false, true);
  b.small_linear(e1, T + "adaln_single.emb.timestep_embedder.linear_2", emb, B, C, C, false, false);
  b.small_linear(emb, T + "adaln_single.linear", t6, B, C, 6 * C, true, false);

  // ---- caption projection: Linear -> GELU(tanh) -> Linear on [B*Lc, caption_channels]
  bf16* cp1 = b.buf(Mc, C);
  bf16* cproj = b.buf(Mc, C);
  {
    Epilogue e;
    e.bias = b.f32(T + "caption_projection.linear_1.bias");
    e.act = kActGeluTanh;
    e.out = cp1;
    e.ld_out = C;
    b.linear(h->ctx_bf16, Mc, a.caption_channels, a.caption_channels, b.lin(T + "caption_projection.linear_1.weight"), C,
             e);
    Epilogue e2;
    e2.bias = b.f32(T + "caption_projection.linear_2.bias");
    e2.out = cproj;
    e2.ld_out = C;
    b.linear(cp1, Mc, C, C, b.lin(T + "caption_projection.linear_2.weight"), C, e2);
  }
  b.rel(cp1);

  // ---- patch embedding + position table: hs = conv(latent) + pos_embed
  bf16* hs = b.buf(M, C);
  {
    const RawW* pe = b.raw(T + "pos_embed.pos_embed");
    if (pe) const_cast<RawW*>(pe)->keep = true;   // replicated into the plan's position table at every gdf_plan
    bf16* pos = nullptr;
    if (pe && !b.dry) {
      if (pe->numel != (int64_t)N * C)
        return b.set_err(fail(GDF_ERR_SHAPE, "pos_embed holds %lld values, this plan needs %d x %d (img_size %d)",
                              (long long)pe->numel, N, C, h->img));
      pos = b.buf(M, C);
      if (pos) {
        cudaError_t ce = launch_replicate_rows_bf16(pe->ptr, pos, (long long)N * C, B, 0);
        if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
        if (ce != cudaSuccess) return b.set_err(fail(GDF_ERR_CUDA, "pos table: %s", cudaGetErrorString(ce)));
      }
    }
    bf16* col = b.buf(M, 64);
    if (!b.dry) {
      bf16* lat = h->latent_nhwc;
      const int L = h->L, cin = a.in_channels;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_patchify(lat, col, B, L, p, cin, 64, rc.stream));
        return 0;
      });
    }
    int npad = 0;
    const bf16* w = b.conv_w(T + "pos_embed.proj.weight", &npad, 64);
    Epilogue e;
    e.bias = b.f32_pad(T + "pos_embed.proj.bias", npad);
    e.n_out = C;
    e.residual = pos;
    e.ld_res = C;
    e.out = hs;
    e.ld_out = C;
    b.linear(col, M, 64, 64, w, npad, e);
    b.rel(col);
    // `pos` stays allocated for the lifetime of the plan (read by every replay)
  }

  // ---- transformer blocks
  float* mod = b.fbuf((long long)6 * B * C);   // [6][B][C]: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
  const long long plane = (long long)B * C;
  for (int i = 0; i < a.num_layers; ++i) {
    const std::string wp = T + "transformer_blocks." + std::to_string(i);
    const std::string fid = "vit-block" + std::to_string(i);
    b.ops->scope = fid;
    const float* table = b.f32(wp + ".scale_shift_table");
    if (!b.dry) {
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_adaln_mod(table, t6, mod, B, 6, C, 6 * C, rc.stream));
        return 0;
      });
    }
    // self attention
    bf16* n1 = b.buf(M, C);
    b.layernorm_mod(hs, n1, M, C, eps, mod + 1 * plane, mod + 0 * plane, N);
    const bf16* wqkv = b.rows_bf16(wp + ".attn1#qkv", {wp + ".attn1.to_q.weight", wp + ".attn1.to_k.weight",
                                                       wp + ".attn1.to_v.weight"}, nullptr);
    const float* bqkv = nullptr;
    {
      const std::string key = wp + ".attn1#qkv_bias";
      auto it = h->packed.find(key);
      if (it != h->packed.end()) bqkv = static_cast<const float*>(it->second);
      else {
        const float* bq = b.f32(wp + ".attn1.to_q.bias");
        const float* bk = b.f32(wp + ".attn1.to_k.bias");
        const float* bv = b.f32(wp + ".attn1.to_v.bias");
        float* d = static_cast<float*>(b.dev_alloc((size_t)3 * C * 4));
        if (d && bq && bk && bv) {
          cudaMemcpy(d, bq, (size_t)C * 4, cudaMemcpyDeviceToDevice);
          cudaMemcpy(d + C, bk, (size_t)C * 4, cudaMemcpyDeviceToDevice);
          cudaMemcpy(d + 2 * C, bv, (size_t)C * 4, cudaMemcpyDeviceToDevice);
        }
        h->packed[key] = d;
        bqkv = d;
      }
    }
    bf16* qkv = b.buf(M, 3 * C);
    {
      Epilogue e;
      e.bias = bqkv;
      e.out = qkv;
      e.ld_out = 3 * C;
      Caps caps;
      caps.add(b.site(fid + "-self-q", C, g, g), 0, C);
      caps.add(b.site(fid + "-self-k", C, g, g), C, 2 * C);
      caps.add(b.site(fid + "-self-v", C, g, g), 2 * C, 3 * C);
      b.linear(n1, M, C, C, wqkv, 3 * C, e, caps);
    }
    b.rel(n1);
    bf16* ao = b.buf(M, C);
    // attention-probability maps (AttnStoreProcessor on attn1 / attn2 of every block, place 'up':
    // feature/components/attention.py:583-593): the materialising kernel for the modules whose map is requested
    if (b.wants(fid + "-self-map") || b.wants("#attnmean:" + fid + "-self"))
      b.attention_probs(fid, "self", qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, 0, ao, C, B, heads, N, N, scale, hd);
    else
      b.attention_bias(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, ao, C, B, heads, N, N, scale, hd, false);
    b.rel(qkv);
    bf16* hs1 = b.buf(M, C);
    {
      Epilogue e;
      e.bias = b.f32(wp + ".attn1.to_out.0.bias");
      e.col_scale = mod + 2 * plane;   // gate_msa
      e.rows_per_batch = N;
      e.residual = hs;
      e.ld_res = C;
      e.out = hs1;
      e.ld_out = C;
      b.linear(ao, M, C, C, b.lin(wp + ".attn1.to_out.0.weight"), C, e);
    }
    b.rel(ao);
    b.rel(hs);
    // cross attention on the un-normalised hidden states (attention.py:539-542)
    bf16* q2 = b.buf(M, C);
    {
      Epilogue e;
      e.bias = b.f32(wp + ".attn2.to_q.bias");
      e.out = q2;
      e.ld_out = C;
      Caps caps;
      caps.add(b.site(fid + "-cross-q", C, g, g), 0, C);
      b.linear(hs1, M, C, C, b.lin(wp + ".attn2.to_q.weight"), C, e, caps);
    }
    const bf16* wkv = b.rows_bf16(wp + ".attn2#kv", {wp + ".attn2.to_k.weight", wp + ".attn2.to_v.weight"}, nullptr);
    const float* bkv = nullptr;
    {
      const std::string key = wp + ".attn2#kv_bias";
      auto it = h->packed.find(key);
      if (it != h->packed.end()) bkv = static_cast<const float*>(it->second);
      else {
        const float* bk = b.f32(wp + ".attn2.to_k.bias");
        const float* bv = b.f32(wp + ".attn2.to_v.bias");
        float* d = static_cast<float*>(b.dev_alloc((size_t)2 * C * 4));
        if (d && bk && bv) {
          cudaMemcpy(d, bk, (size_t)C * 4, cudaMemcpyDeviceToDevice);
          cudaMemcpy(d + C, bv, (size_t)C * 4, cudaMemcpyDeviceToDevice);
        }
        h->packed[key] = d;
        bkv = d;
      }
    }
    bf16* kv = b.buf(Mc, 2 * C);
    {
      Epilogue e;
      e.bias = bkv;
      e.out = kv;
      e.ld_out = 2 * C;
      b.linear(cproj, Mc, C, C, wkv, 2 * C, e);
    }
    bf16* ao2 = b.buf(M, C);
    if (b.wants(fid + "-cross-map") || b.wants("#attnmean:" + fid + "-cross"))
      b.attention_probs(fid, "cross", q2, C, kv, 2 * C, kv + C, 2 * C, 0, ao2, C, B, heads, N, Lc, scale, hd, true);
    else
      b.attention_bias(q2, C, kv, 2 * C, kv + C, 2 * C, ao2, C, B, heads, N, Lc, scale, hd, true);
    b.rel(q2);
    b.rel(kv);
    bf16* hs2 = b.buf(M, C);
    {
      Epilogue e;
      e.bias = b.f32(wp + ".attn2.to_out.0.bias");
      e.residual = hs1;
      e.ld_res = C;
      e.out = hs2;
      e.ld_out = C;
      b.linear(ao2, M, C, C, b.lin(wp + ".attn2.to_out.0.weight"), C, e);
    }
    b.rel(ao2);
    b.rel(hs1);
    // feed-forward: norm2 + mlp modulation, GELU(tanh) (attention.py:570-583, 1249-1258)
    bf16* n2 = b.buf(M, C);
    b.layernorm_mod(hs2, n2, M, C, eps, mod + 4 * plane, mod + 3 * plane, N);
    const int inner = 4 * C;
    bf16* ffi = b.buf(M, inner);
    {
      Epilogue e;
      e.act = kActGeluTanh;
      e.bias = b.f32(wp + ".ff.net.0.proj.bias");
      e.out = ffi;
      e.ld_out = inner;
      Caps caps;
      caps.add(b.site(fid + "-ffn-inner", inner, g, g), 0, inner);
      b.linear(n2, M, C, C, b.lin(wp + ".ff.net.0.proj.weight"), inner, e, caps);
    }
    b.rel(n2);
    bf16* hs3 = b.buf(M, C);
    {
      Epilogue e;
      e.bias = b.f32(wp + ".ff.net.2.bias");
      e.col_scale = mod + 5 * plane;   // gate_mlp
      e.rows_per_batch = N;
      e.residual = hs2;
      e.ld_res = C;
      e.out = hs3;
      e.ld_out = C;
      Caps caps;
      caps.add(b.site(fid + "-out", C, g, g), 0, C);
      b.linear(ffi, M, inner, inner, b.lin(wp + ".ff.net.2.weight"), C, e, caps);
    }
    b.rel(ffi);
    b.rel(hs2);
    hs = hs3;
  }
  b.rel(cproj);

  // ---- output: norm_out + (scale_shift_table + embedded_timestep) modulation -> proj_out -> unpatchify
  {
    float* mod2 = b.fbuf((long long)2 * B * C);   // [2][B][C]: shift, scale
    const float* table = b.f32(T + "scale_shift_table");
    if (!b.dry) {
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_adaln_mod(table, emb, mod2, B, 2, C, C, rc.stream));
        return 0;
      });
    }
    bf16* nf = b.buf(M, C);
    b.layernorm_mod(hs, nf, M, C, eps, mod2 + plane, mod2, N);
    b.rel(hs);
    const int pout = p * p * a.out_channels;
    const int npad = (pout + 15) / 16 * 16;
    std::vector<int> idx(npad);
    for (int r = 0; r < npad; ++r) idx[r] = r < pout ? r : -1;
    const bf16* w = b.rows_bf16(T + "proj_out#pad", {T + "proj_out.weight"}, &idx);
    float* o = b.fbuf(M * npad);
    Epilogue e;
    e.bias = b.f32_pad(T + "proj_out.bias", npad);
    e.n_out = pout;
    e.out_f32 = o;
    e.ld_out_f32 = npad;
    b.linear(nf, M, C, C, w, npad, e);
    b.rel(nf);
    if (!b.dry) {
      const int oc = a.out_channels;
      if (npad != pout) return b.set_err(fail(GDF_ERR_UNSUPPORTED, "DiT: patch^2 * out_channels must be a multiple of 16"));
      b.ops->push_back([=](const RunCtx& rc) -> int {
        if (rc.noise_pred_out) OP_CUDA(launch_unpatchify(o, rc.noise_pred_out, B, g, p, oc, rc.stream));
        return 0;
      });
    }
  }
  return b.err;
}

// ----------------------------------------------------------------------------------------- Flux MMDiT
// FluxTransformer2DModel.forward (transformer_flux.py:414-604) with every gather call site of the reference's Flux
// branch (feature_extractor.py:98-123): FluxAttnProcessor2_0 q / k / v / attn-out (attention_processor.py:2280-2289,
// 2355-2361), FluxTransformerBlock norm-out / out (transformer_flux.py:200-211; both store norm_hidden_states - a
// quirk of the reference that is kept), FeedForward inner (attention.py:1249-1258), FluxSingleTransformerBlock out
// (:107-108, image rows only).
// The op list is emitted sample by sample: the token sequence of one image is the concatenation [text | image]
// (S = ctx_len + N rows of one joint buffer), so every projection is a plain row-range GEMM, the text and image
// streams of a double block are two launches with their own weights on the two row ranges, and capture slots
// (image rows only) are addressed by a per-sample offset. Weights (24 GB bf16 at full size) are re-read per sample,
// which costs < 3 % of a sample's tensor-pipe time.
static int build_flux(Builder& b) {
  gdf_handle_s* h = b.h;
  const gdf_flux_arch& a = h->fa;
  const int B = h->B, heads = a.num_heads, hd = a.head_dim;
  const int C = heads * hd, g = h->L / 2, N = g * g, Lt = h->ctx_len, S = Lt + N;
  const int J = a.joint_attention_dim, P = a.pooled_projection_dim, cin = a.in_channels, lc = cin / 4;
  const float eps = 1e-6f;
  const float scale = 1.f / sqrtf((float)hd);
  const std::string T = "transformer.";
  b.ops = &h->unet_ops;
  b.ops->scope = "denoiser-head";
  b.gn_fuse = false;
  b.gn_cap = 0;
  h->unet_in_cap = -1;
  if (C % 64 != 0 || hd % 8 != 0 || hd > 160 || cin % 8 != 0 || cin != 4 * h->va.latent_channels || J % 8 != 0 ||
      Lt % 8 != 0)
    return b.set_err(fail(GDF_ERR_UNSUPPORTED, "Flux: hidden %d / head_dim %d / in_channels %d / ctx %d x %d unsupported",
                          C, hd, cin, Lt, J));

  // ---- conditioning (all B rows at once, fp32): temb = time(sigma * 1000) + guidance(g * 1000) + text(pooled)
  // [diffusers embeddings.CombinedTimestepGuidanceTextProjEmbeddings, un-vendored; called at transformer_flux.py:463-467]
  float* tsin = b.fbuf((long long)B * 256);
  float* gsin = b.fbuf((long long)B * 256);
  float* e1 = b.fbuf((long long)B * C);
  float* et = b.fbuf((long long)B * C);
  float* eg = b.fbuf((long long)B * C);
  float* ep = b.fbuf((long long)B * C);
  float* temb = b.fbuf((long long)B * C);
  float* pool_in = b.fbuf((long long)B * P);
  if (!b.dry) {
    float* t_dev = h->t_dev;
    float* g_dev = h->g_dev;
    const bool guid = a.guidance_embeds != 0;
    gdf_handle_s* hh = h;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_timestep_embedding(t_dev, tsin, B, 256, rc.stream));
      if (guid) OP_CUDA(launch_timestep_embedding(g_dev, gsin, B, 256, rc.stream));
      OP_CUDA(cudaMemcpyAsync(pool_in, hh->pooled, (size_t)B * P * 4, cudaMemcpyDeviceToDevice, rc.stream));
      return 0;
    });
  }
  b.small_linear(tsin, T + "time_text_embed.timestep_embedder.linear_1", e1, B, 256, C, false, true);
  b.small_linear(e1, T + "time_text_embed.timestep_embedder.linear_2", et, B, C, C, false, false);
  if (a.guidance_embeds) {
    b.small_linear(gsin, T + "time_text_embed.guidance_embedder.linear_1", e1, B, 256, C, false, true);
    b.small_linear(e1, T + "time_text_embed.guidance_embedder.linear_2", eg, B, C, C, false, false);
  }
  b.small_linear(pool_in, T + "time_text_embed.text_embedder.linear_1", e1, B, P, C, false, true);
  b.small_linear(e1, T + "time_text_embed.text_embedder.linear_2", ep, B, C, C, false, false);
  if (!b.dry) {
    const bool guid = a.guidance_embeds != 0;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_sum3_f32(temb, et, ep, guid ? eg : nullptr, B * C, rc.stream));
      return 0;
    });
  }
  // modulation vectors of every block: Linear(SiLU(temb)) [diffusers normalization.AdaLayerNormZero (6C: shift_msa,
  // scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp), AdaLayerNormZeroSingle (3C: shift, scale, gate),
  // AdaLayerNormContinuous (2C: scale, shift), un-vendored]
  std::vector<float*> mod_img(a.num_layers), mod_txt(a.num_layers), mod_s(a.num_single_layers);
  for (int i = 0; i < a.num_layers; ++i) {
    const std::string wp = T + "transformer_blocks." + std::to_string(i);
    mod_img[i] = b.fbuf((long long)B * 6 * C);
    mod_txt[i] = b.fbuf((long long)B * 6 * C);
    b.small_linear(temb, wp + ".norm1.linear", mod_img[i], B, C, 6 * C, true, false);
    b.small_linear(temb, wp + ".norm1_context.linear", mod_txt[i], B, C, 6 * C, true, false);
  }
  for (int i = 0; i < a.num_single_layers; ++i) {
    mod_s[i] = b.fbuf((long long)B * 3 * C);
    b.small_linear(temb, T + "single_transformer_blocks." + std::to_string(i) + ".norm.linear", mod_s[i], B, C, 3 * C,
                   true, false);
  }
  float* mod_f = b.fbuf((long long)B * 2 * C);
  b.small_linear(temb, T + "norm_out.linear", mod_f, B, C, 2 * C, true, false);

  // ---- packed latents: (B, lc, L, L) -> (B, N, 4 lc), column = c * 4 + py * 2 + px (_pack_latents)
  bf16* col = b.buf((long long)B * N, cin);
  if (!b.dry) {
    bf16* lat = h->latent_nhwc;
    const int L = h->L;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_patchify(lat, col, B, L, 2, lc, cin, rc.stream, 1));
      return 0;
    });
  }

  // capture slots: registered once (execution order of sample 0), addressed per sample
  std::unordered_map<std::string, int64_t> slot_of;
  auto slot = [&](const std::string& id, int Cc, int s) -> int64_t {
    auto it = slot_of.find(id);
    int64_t base;
    if (it == slot_of.end()) {
      base = b.site(id, Cc, g, g);
      slot_of[id] = base;
    } else {
      base = it->second;
    }
    return base < 0 ? -1 : base + (int64_t)s * N * Cc * 2;
  };
  // fp16 copy of bf16 rows into a capture slot (tensors no GEMM epilogue produces)
  auto capture_rows = [&](const bf16* src, int ld_src, int64_t off, int Cc) {
    if (b.dry || b.err || off < 0) return;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_copy_rows_bf16_f16(src, ld_src, reinterpret_cast<__half*>(rc.arena + off), Cc, N, Cc, rc.stream));
      return 0;
    });
  };
  // Joint attention of sample `s` over the [text | image] rows of qkv. With an attention-probability map (or its head
  // mean for the aggregated `attn` feature) requested for this block - FluxAttnStoreProcessor,
  // feature/components/attention.py:402-527 - the image-query rows run the materialising kernel, which writes
  // P[:, :, image rows, text keys] (`cross-map`) and P[:, :, image rows, image keys] (`self-map`) separately; the text
  // rows (whose probabilities the reference does not store) stay on the flash kernel.
  std::unordered_map<std::string, int64_t> map_slot_of;
  auto map_slot = [&](const std::string& id, int Nk_, int s) -> int64_t {
    auto it = map_slot_of.find(id);
    int64_t base;
    if (it == map_slot_of.end()) {
      base = b.site(id, id[0] == '#' ? 1 : heads, N, Nk_);
      map_slot_of[id] = base;
    } else {
      base = it->second;
    }
    return base < 0 ? -1 : base + (int64_t)s * (id[0] == '#' ? 1 : heads) * N * Nk_ * 2;
  };
  auto joint_attention = [&](const std::string& fid, int s, bf16* qkv, bf16* out, int ld_out) {
    const bool maps = b.wants(fid + "-cross-map") || b.wants(fid + "-self-map") ||
                      b.wants("#attnmean:" + fid + "-cross") || b.wants("#attnmean:" + fid + "-self");
    if (!maps) {
      b.attention_bias(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, out, ld_out, 1, heads, S, S, scale, hd, false);
      return;
    }
    if (b.dry || b.err) return;
    // registration order = the reference's gather order: cross-map, self-map (then the internal head means)
    const int64_t cross_off = map_slot(fid + "-cross-map", Lt, s);
    const int64_t self_off = map_slot(fid + "-self-map", N, s);
    const int64_t cmean_off = map_slot("#attnmean:" + fid + "-cross", Lt, s);
    const int64_t smean_off = map_slot("#attnmean:" + fid + "-self", N, s);
    __half* cross_scr = (cross_off < 0 && cmean_off >= 0) ? reinterpret_cast<__half*>(b.buf((long long)heads * N, Lt)) : nullptr;
    __half* self_scr = (self_off < 0 && smean_off >= 0) ? reinterpret_cast<__half*>(b.buf((long long)heads * N, N)) : nullptr;
    b.attention_bias(qkv, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, out, ld_out, 1, heads, Lt, S, scale, hd, false);
    b.ops->tag(kKindAttention, 4.0 * heads * (double)N * (double)S * hd,
               "attention-probs (joint, image rows) heads=" + std::to_string(heads) + " d=" + std::to_string(hd));
    const bf16* q_img = qkv + (long long)Lt * 3 * C;
    bf16* o_img = out + (long long)Lt * ld_out;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      __half* Pc = cross_off >= 0 ? reinterpret_cast<__half*>(rc.arena + cross_off) : cross_scr;
      __half* Ps = self_off >= 0 ? reinterpret_cast<__half*>(rc.arena + self_off) : self_scr;
      OP_CUDA(launch_attention_probs(q_img, 3 * C, qkv + C, 3 * C, qkv + 2 * C, 3 * C, 0, o_img, ld_out, Ps, 1, heads, N, S,
                                     hd, scale, rc.stream, nullptr, Pc, Lt));
      if (cmean_off >= 0)
        OP_CUDA(launch_head_mean(Pc, reinterpret_cast<__half*>(rc.arena + cmean_off), 1, heads, (long long)N * Lt, rc.stream));
      if (smean_off >= 0)
        OP_CUDA(launch_head_mean(Ps, reinterpret_cast<__half*>(rc.arena + smean_off), 1, heads, (long long)N * N, rc.stream));
      return 0;
    });
    if (cross_scr) b.rel(reinterpret_cast<bf16*>(cross_scr));
    if (self_scr) b.rel(reinterpret_cast<bf16*>(self_scr));
  };
  auto qk_norm_rope = [&](bf16* qkv, const float* wq_a, const float* wk_a, const float* wq_b, const float* wk_b,
                          int rows_a) {
    if (b.dry || b.err) return;
    gdf_handle_s* hh = h;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      OP_CUDA(launch_qk_rmsnorm_rope(qkv, 3 * C, S, heads, hd, C, wq_a, wk_a, wq_b, wk_b, rows_a, hh->rope_cos,
                                     hh->rope_sin, eps, rc.stream));
      return 0;
    });
  };

  for (int s = 0; s < B; ++s) {
    // ---- embedders: joint hidden state [S, C] = [context_embedder(ctx) | x_embedder(packed latents)]
    bf16* hs = b.buf(S, C);
    {
      Epilogue e;
      e.bias = b.f32(T + "context_embedder.bias");
      e.out = hs;
      e.ld_out = C;
      b.linear(h->ctx_bf16 + (long long)s * Lt * J, Lt, J, J, b.lin(T + "context_embedder.weight"), C, e);
      Epilogue e2;
      e2.bias = b.f32(T + "x_embedder.bias");
      e2.out = hs + (long long)Lt * C;
      e2.ld_out = C;
      b.linear(col + (long long)s * N * cin, N, cin, cin, b.lin(T + "x_embedder.weight"), C, e2);
    }

    // ---- double-stream blocks
    for (int i = 0; i < a.num_layers; ++i) {
      const std::string wp = T + "transformer_blocks." + std::to_string(i);
      const std::string fid = "vit-block" + std::to_string(i);
      b.ops->scope = fid;
      const float* mi = mod_img[i] + (long long)s * 6 * C;
      const float* mt = mod_txt[i] + (long long)s * 6 * C;
      bf16* hs_t = hs;
      bf16* hs_i = hs + (long long)Lt * C;
      bf16* nj = b.buf(S, C);
      b.layernorm_mod(hs_t, nj, Lt, C, eps, mt + C, mt, Lt);
      b.layernorm_mod(hs_i, nj + (long long)Lt * C, N, C, eps, mi + C, mi, N);
      const bf16* wqkv_i = b.rows_bf16(wp + ".attn#qkv", {wp + ".attn.to_q.weight", wp + ".attn.to_k.weight",
                                                          wp + ".attn.to_v.weight"}, nullptr);
      const float* bqkv_i = b.f32_cat(wp + ".attn#qkv_bias", {wp + ".attn.to_q.bias", wp + ".attn.to_k.bias",
                                                              wp + ".attn.to_v.bias"});
      const bf16* wqkv_t = b.rows_bf16(wp + ".attn#add_qkv", {wp + ".attn.add_q_proj.weight",
                                                              wp + ".attn.add_k_proj.weight",
                                                              wp + ".attn.add_v_proj.weight"}, nullptr);
      const float* bqkv_t = b.f32_cat(wp + ".attn#add_qkv_bias", {wp + ".attn.add_q_proj.bias",
                                                                  wp + ".attn.add_k_proj.bias",
                                                                  wp + ".attn.add_v_proj.bias"});
      bf16* qkv = b.buf(S, 3 * C);
      {
        Epilogue e;
        e.bias = bqkv_t;
        e.out = qkv;
        e.ld_out = 3 * C;
        b.linear(nj, Lt, C, C, wqkv_t, 3 * C, e);
        Epilogue e2;
        e2.bias = bqkv_i;
        e2.out = qkv + (long long)Lt * 3 * C;
        e2.ld_out = 3 * C;
        Caps caps;
        caps.add(slot(fid + "-q", C, s), 0, C);
        caps.add(slot(fid + "-k", C, s), C, 2 * C);
        caps.add(slot(fid + "-v", C, s), 2 * C, 3 * C);
        b.linear(nj + (long long)Lt * C, N, C, C, wqkv_i, 3 * C, e2, caps);
      }
      b.rel(nj);
      qk_norm_rope(qkv, b.f32(wp + ".attn.norm_added_q.weight"), b.f32(wp + ".attn.norm_added_k.weight"),
                   b.f32(wp + ".attn.norm_q.weight"), b.f32(wp + ".attn.norm_k.weight"), Lt);
      bf16* ao = b.buf(S, C);
      joint_attention(fid, s, qkv, ao, C);
      b.rel(qkv);
      bf16* hs2 = b.buf(S, C);
      {
        Epilogue e;   // image stream: hs + gate_msa * to_out(attn); attn-out captured before the gate
        e.bias = b.f32(wp + ".attn.to_out.0.bias");
        e.col_scale = mi + 2 * C;
        e.rows_per_batch = N;
        e.residual = hs_i;
        e.ld_res = C;
        e.out = hs2 + (long long)Lt * C;
        e.ld_out = C;
        Caps caps;
        caps.pre = slot(fid + "-attn-out", C, s);
        b.linear(ao + (long long)Lt * C, N, C, C, b.lin(wp + ".attn.to_out.0.weight"), C, e, caps);
        Epilogue e2;  // text stream
        e2.bias = b.f32(wp + ".attn.to_add_out.bias");
        e2.col_scale = mt + 2 * C;
        e2.rows_per_batch = Lt;
        e2.residual = hs_t;
        e2.ld_res = C;
        e2.out = hs2;
        e2.ld_out = C;
        b.linear(ao, Lt, C, C, b.lin(wp + ".attn.to_add_out.weight"), C, e2);
      }
      b.rel(ao);
      b.rel(hs);
      bf16* n2 = b.buf(S, C);
      b.layernorm_mod(hs2 + (long long)Lt * C, n2 + (long long)Lt * C, N, C, eps, mi + 4 * C, mi + 3 * C, N);
      b.layernorm_mod(hs2, n2, Lt, C, eps, mt + 4 * C, mt + 3 * C, Lt);
      capture_rows(n2 + (long long)Lt * C, C, slot(fid + "-norm-out", C, s), C);
      const int inner = 4 * C;
      bf16* ffi = b.buf(S, inner);
      {
        Epilogue e;
        e.act = kActGeluTanh;
        e.bias = b.f32(wp + ".ff.net.0.proj.bias");
        e.out = ffi + (long long)Lt * inner;
        e.ld_out = inner;
        Caps caps;
        caps.add(slot(fid + "-ffn-inner", inner, s), 0, inner);
        b.linear(n2 + (long long)Lt * C, N, C, C, b.lin(wp + ".ff.net.0.proj.weight"), inner, e, caps);
        Epilogue e2;
        e2.act = kActGeluTanh;
        e2.bias = b.f32(wp + ".ff_context.net.0.proj.bias");
        e2.out = ffi;
        e2.ld_out = inner;
        b.linear(n2, Lt, C, C, b.lin(wp + ".ff_context.net.0.proj.weight"), inner, e2);
      }
      capture_rows(n2 + (long long)Lt * C, C, slot(fid + "-out", C, s), C);   // transformer_flux.py:210-211
      b.rel(n2);
      bf16* hs3 = b.buf(S, C);
      {
        Epilogue e;
        e.bias = b.f32(wp + ".ff.net.2.bias");
        e.col_scale = mi + 5 * C;
        e.rows_per_batch = N;
        e.residual = hs2 + (long long)Lt * C;
        e.ld_res = C;
        e.out = hs3 + (long long)Lt * C;
        e.ld_out = C;
        b.linear(ffi + (long long)Lt * inner, N, inner, inner, b.lin(wp + ".ff.net.2.weight"), C, e);
        Epilogue e2;
        e2.bias = b.f32(wp + ".ff_context.net.2.bias");
        e2.col_scale = mt + 5 * C;
        e2.rows_per_batch = Lt;
        e2.residual = hs2;
        e2.ld_res = C;
        e2.out = hs3;
        e2.ld_out = C;
        b.linear(ffi, Lt, inner, inner, b.lin(wp + ".ff_context.net.2.weight"), C, e2);
      }
      b.rel(ffi);
      b.rel(hs2);
      hs = hs3;
    }

    // ---- single-stream blocks on the joint sequence (transformer_flux.py:539-541 cat([text, image]))
    for (int i = 0; i < a.num_single_layers; ++i) {
      const std::string wp = T + "single_transformer_blocks." + std::to_string(i);
      const std::string fid = "vit-block" + std::to_string(a.num_layers + i);
      b.ops->scope = fid;
      const float* ms = mod_s[i] + (long long)s * 3 * C;
      bf16* nj = b.buf(S, C);
      b.layernorm_mod(hs, nj, S, C, eps, ms + C, ms, S);
      const bf16* wqkv = b.rows_bf16(wp + ".attn#qkv", {wp + ".attn.to_q.weight", wp + ".attn.to_k.weight",
                                                        wp + ".attn.to_v.weight"}, nullptr);
      const float* bqkv = b.f32_cat(wp + ".attn#qkv_bias", {wp + ".attn.to_q.bias", wp + ".attn.to_k.bias",
                                                            wp + ".attn.to_v.bias"});
      bf16* qkv = b.buf(S, 3 * C);
      bf16* cat = b.buf(S, 5 * C);   // [attention output | GELU(proj_mlp)] = the operand of proj_out
      {
        Epilogue e;
        e.bias = bqkv;
        e.out = qkv;
        e.ld_out = 3 * C;
        b.linear(nj, Lt, C, C, wqkv, 3 * C, e);
        Epilogue e2;
        e2.bias = bqkv;
        e2.out = qkv + (long long)Lt * 3 * C;
        e2.ld_out = 3 * C;
        Caps caps;
        caps.add(slot(fid + "-q", C, s), 0, C);
        caps.add(slot(fid + "-k", C, s), C, 2 * C);
        caps.add(slot(fid + "-v", C, s), 2 * C, 3 * C);
        b.linear(nj + (long long)Lt * C, N, C, C, wqkv, 3 * C, e2, caps);
        Epilogue e3;
        e3.act = kActGeluTanh;
        e3.bias = b.f32(wp + ".proj_mlp.bias");
        e3.out = cat + C;
        e3.ld_out = 5 * C;
        b.linear(nj, S, C, C, b.lin(wp + ".proj_mlp.weight"), 4 * C, e3);
      }
      b.rel(nj);
      const float* nq = b.f32(wp + ".attn.norm_q.weight");
      const float* nk = b.f32(wp + ".attn.norm_k.weight");
      qk_norm_rope(qkv, nq, nk, nq, nk, 0);
      joint_attention(fid, s, qkv, cat, 5 * C);
      b.rel(qkv);
      capture_rows(cat + (long long)Lt * 5 * C, 5 * C, slot(fid + "-attn-out", C, s), C);
      bf16* hs2 = b.buf(S, C);
      {
        const bf16* wo = b.lin(wp + ".proj_out.weight");
        Epilogue e;
        e.bias = b.f32(wp + ".proj_out.bias");
        e.col_scale = ms + 2 * C;
        e.rows_per_batch = Lt;
        e.residual = hs;
        e.ld_res = C;
        e.out = hs2;
        e.ld_out = C;
        b.linear(cat, Lt, 5 * C, 5 * C, wo, C, e);
        Epilogue e2 = e;
        e2.rows_per_batch = N;
        e2.residual = hs + (long long)Lt * C;
        e2.out = hs2 + (long long)Lt * C;
        Caps caps;
        caps.add(slot(fid + "-out", C, s), 0, C);
        b.linear(cat + (long long)Lt * 5 * C, N, 5 * C, 5 * C, wo, C, e2, caps);
      }
      b.rel(cat);
      b.rel(hs);
      hs = hs2;
    }

    // ---- norm_out (AdaLayerNormContinuous: scale first, then shift) + proj_out on the image rows
    {
      const float* mf = mod_f + (long long)s * 2 * C;
      bf16* nf = b.buf(N, C);
      b.layernorm_mod(hs + (long long)Lt * C, nf, N, C, eps, mf, mf + C, N);
      b.rel(hs);
      const int pout = cin;
      const int npad = (pout + 15) / 16 * 16;
      std::vector<int> idx(npad);
      for (int r = 0; r < npad; ++r) idx[r] = r < pout ? r : -1;
      const bf16* w = b.rows_bf16(T + "proj_out#pad", {T + "proj_out.weight"}, &idx);
      float* o = b.fbuf((long long)N * npad);
      Epilogue e;
      e.bias = b.f32_pad(T + "proj_out.bias", npad);
      e.n_out = pout;
      e.out_f32 = o;
      e.ld_out_f32 = npad;
      b.linear(nf, N, C, C, w, npad, e);
      b.rel(nf);
      if (!b.dry) {
        b.ops->push_back([=](const RunCtx& rc) -> int {
          if (rc.noise_pred_out)
            OP_CUDA(cudaMemcpy2DAsync(rc.noise_pred_out + (long long)s * N * pout, (size_t)pout * 4, o, (size_t)npad * 4,
                                      (size_t)pout * 4, N, cudaMemcpyDeviceToDevice, rc.stream));
          return 0;
        });
      }
      b.rel(o);
    }
  }
  b.rel(col);
  return b.err;
}

// ----------------------------------------------------------------------------------------- VAE encoder
// UNetMidBlock2D of the VAE (encoder and decoder): resnet, single-head attention over all pixels (d = ch), resnet.
// mp = "vae.encoder.mid_block" / "vae.decoder.mid_block". cur / cur_sums are replaced by the block's output.
static void emit_vae_mid_block(Builder& b, const std::string& mp, bf16*& cur, float*& cur_sums, int B, int hw, int ch, int G,
                               float eps) {
  const long long M = (long long)B * hw * hw;
  const int N = hw * hw;
    bf16* r0 = b.buf(M, ch);
    Dest d;
    d.out = r0;
    d.ld = ch;
    d.gn_sums = b.gn_fusable(ch, G, (long long)N) ? b.gn_slot(B, G) : nullptr;
    emit_resnet(b, mp + ".resnets.0", "", cur, B, hw, hw, ch, ch, nullptr, 0, G, eps, d, cur_sums);
    b.rel(cur);
    const std::string ap = mp + ".attentions.0";
    bf16* hn = b.buf(M, ch);
    if (d.gn_sums) b.groupnorm_from_sums(r0, hn, ap + ".group_norm", B, N, ch, G, eps, false, d.gn_sums);
    else b.groupnorm(r0, hn, ap + ".group_norm", B, N, ch, G, eps, false);
    // Q | K fused projection
    const bf16* wqk = b.rows_bf16(ap + "#qk", {ap + ".to_q.weight", ap + ".to_k.weight"}, nullptr);
    // The concatenated q | k bias is a CONSTANT of the plan: it lives in weight storage (f32_cat), not in the activation
    // pool. Round 1 filled a pool buffer at plan time and released it after the block: the pool handed it to a later
    // layer, so from the second forward on (the first one still saw the plan-time contents) the VAE's mid-block
    // attention added whatever activations that layer had left there instead of the bias. Found in round 2 through the
    // bit-reproducibility probe (tools/probe_determinism_vae2.py: the latents changed for good after the first UNet
    // forward); the error was small enough (logit shifts of one 512-wide head) to pass the cosine / max-relative bounds.
    const float* bqk = b.f32_cat(ap + "#qk_bias", {ap + ".to_q.bias", ap + ".to_k.bias"});
    bf16* qk = b.buf(M, 2 * ch);
    {
      Epilogue e;
      e.bias = bqk;
      e.out = qk;
      e.ld_out = 2 * ch;
      b.linear(hn, M, ch, ch, wqk, 2 * ch, e);
    }
    // V^T[b] = Wv hn[b]^T + bv[:, None]  -> [B, ch, N]
    bf16* vt = b.buf((long long)B * ch, N);
    {
      Epilogue e;
      e.bias_m = b.f32(ap + ".to_v.bias");
      e.out = vt;
      e.ld_out = N;
      e.out_batch_stride = (long long)ch * N;
      b.linear(b.lin(ap + ".to_v.weight"), ch, ch, ch, hn, N, e, Caps(), B, 0, (long long)N * ch, ch);
    }
    b.rel(hn);
    // S[b] = Q[b] K[b]^T / sqrt(ch)
    bf16* S = b.buf((long long)B * N, N);
    {
      Epilogue e;
      e.alpha = 1.f / sqrtf((float)ch);
      e.out = S;
      e.ld_out = N;
      e.out_batch_stride = (long long)N * N;
      b.linear(qk, N, ch, 2 * ch, qk + ch, N, e, Caps(), B, (long long)N * 2 * ch, (long long)N * 2 * ch,
               2 * ch);
    }
    b.rel(qk);
    if (!b.dry) {
      const long long rows = (long long)B * N;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_softmax_rows(S, rows, N, N, rc.stream));
        return 0;
      });
    }
    // O[b] = P[b] V[b]
    bf16* o = b.buf(M, ch);
    {
      Epilogue e;
      e.out = o;
      e.ld_out = ch;
      e.out_batch_stride = (long long)N * ch;
      b.linear(S, N, N, N, vt, ch, e, Caps(), B, (long long)N * N, (long long)ch * N, N);
    }
    b.rel(S);
    b.rel(vt);
    bf16* ao = b.buf(M, ch);
    {
      Epilogue e;
      e.bias = b.f32(ap + ".to_out.0.bias");
      e.residual = r0;
      e.ld_res = ch;
      e.out = ao;
      e.ld_out = ch;
      cur_sums = b.gn_fusable(ch, G, (long long)N) ? b.gn_slot(B, G) : nullptr;
      Builder::want_gn_stats(e, cur_sums, ch, G, (long long)N);
      b.linear(o, M, ch, ch, b.lin(ap + ".to_out.0.weight"), ch, e);
    }
    b.rel(o);
    b.rel(r0);
    bf16* r1 = b.buf(M, ch);
    Dest d1;
    d1.out = r1;
    d1.ld = ch;
    d1.gn_sums = b.gn_fusable(ch, G, (long long)N) ? b.gn_slot(B, G) : nullptr;
    emit_resnet(b, mp + ".resnets.1", "", ao, B, hw, hw, ch, ch, nullptr, 0, G, eps, d1, cur_sums);
    b.rel(ao);
    cur = r1;
    cur_sums = d1.gn_sums;
}

// ------------------------------------------------------------------------------------------ VAE decoder (vae-out)
// z = c_latent * latents + c_model * model_out (scheduler.step + the 1 / scaling_factor of diffusion_feature.py:478-483,
// both linear in the two tensors: schedulers.step_coeffs), then AutoencoderKL.post_quant_conv (1x1, fp32) -> bf16 NHWC.
__global__ void decode_prep_kernel(const float* __restrict__ latents, const float* __restrict__ model_out, float c_latent,
                                   float c_model, const float* __restrict__ pq_w, const float* __restrict__ pq_b,
                                   bf16* __restrict__ z, int B, int HW, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // pixel
  if (i >= (long long)B * HW) return;
  const long long bi = i / HW, px = i - bi * HW;
  float v[16];
  for (int c = 0; c < C; ++c) {
    const long long src = (bi * C + c) * HW + px;
    v[c] = c_latent * latents[src] + (model_out ? c_model * model_out[src] : 0.f);
  }
  for (int o = 0; o < C; ++o) {
    float acc = v[o];
    if (pq_w) {
      acc = pq_b[o];
      for (int c = 0; c < C; ++c) acc = fmaf(pq_w[o * C + c], v[c], acc);
    }
    z[i * C + o] = __float2bfloat16(acc);
  }
}

// [diffusers autoencoders/vae.Decoder, un-vendored]: conv_in, mid block, up blocks over reversed(block_out_channels)
// with layers_per_block + 1 resnets and nearest-x2 + conv upsamplers, GroupNorm + SiLU + conv_out -> fp32 NHWC image.
static int build_vae_decoder(Builder& b) {
  gdf_handle_s* h = b.h;
  const gdf_vae_arch& a = h->va;
  const int B = h->B, G = a.norm_num_groups, lat = a.latent_channels, nl = a.num_levels;
  const float eps = a.norm_eps;
  const std::string V = "vae.decoder.";
  if (9 * lat > 64 || lat > 16)
    return fail(GDF_ERR_UNSUPPORTED, "VAE decoder: %d latent channels (conv_in runs as a K = 64 im2col GEMM: <= 7)", lat);
  b.ops = &h->dec_ops;
  b.ops->scope = "vae-decoder";
  int hw = h->L;
  int ch = a.block_out_channels[nl - 1];
  b.gn_begin(B, G, 64);
  const bool has_pq = h->raw.count("vae.post_quant_conv.weight") != 0;
  const float* pq_w = has_pq ? b.f32("vae.post_quant_conv.weight") : nullptr;
  const float* pq_b = has_pq ? b.f32("vae.post_quant_conv.bias") : nullptr;
  bf16* z = b.buf((long long)B * hw * hw, lat);
  bf16* col = b.buf((long long)B * hw * hw, 64);
  if (!b.dry) {
    const int HW = hw * hw, S = hw;
    b.ops->push_back([=](const RunCtx& rc) -> int {
      const long long n = (long long)B * HW;
      decode_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, rc.stream>>>(rc.dec_latents, rc.dec_model_out, rc.dec_c_latent,
                                                                          rc.dec_c_model, pq_w, pq_b, z, B, HW, lat);
      OP_CUDA(cudaGetLastError());
      OP_CUDA(launch_im2col_small(nullptr, z, col, B, S, S, lat, rc.stream));
      return 0;
    });
  }
  float* cur_sums = b.gn_fusable(ch, G, (long long)hw * hw) ? b.gn_slot(B, G) : nullptr;
  bf16* cur = b.buf((long long)B * hw * hw, ch);
  {
    int npad = 0;
    const bf16* w = b.conv_w(V + "conv_in.weight", &npad, 64);
    Epilogue e;
    e.bias = b.f32_pad(V + "conv_in.bias", npad);
    e.n_out = ch;
    e.out = cur;
    e.ld_out = ch;
    Builder::want_gn_stats(e, cur_sums, ch, G, (long long)hw * hw);
    b.linear(col, (long long)B * hw * hw, 64, 64, w, npad, e);
  }
  b.rel(col);
  b.rel(z);
  emit_vae_mid_block(b, "vae.decoder.mid_block", cur, cur_sums, B, hw, ch, G, eps);
  for (int i = 0; i < nl; ++i) {
    const int cout = a.block_out_channels[nl - 1 - i];
    const std::string up = V + "up_blocks." + std::to_string(i);
    for (int j = 0; j < a.layers_per_block + 1; ++j) {
      bf16* o = b.buf((long long)B * hw * hw, cout);
      Dest d;
      d.out = o;
      d.ld = cout;
      // the output feeds a GroupNorm (next resnet / conv_norm_out) unless the upsampler follows
      const bool before_up = (j == a.layers_per_block) && (i != nl - 1);
      d.gn_sums = (!before_up && b.gn_fusable(cout, G, (long long)hw * hw)) ? b.gn_slot(B, G) : nullptr;
      emit_resnet(b, up + ".resnets." + std::to_string(j), "", cur, B, hw, hw, ch, cout, nullptr, 0, G, eps, d, cur_sums);
      b.rel(cur);
      cur = o;
      cur_sums = d.gn_sums;
      ch = cout;
    }
    if (i != nl - 1) {   // Upsample2D: nearest x2 + conv3x3 (upsampling.py:176-193)
      bf16* big = b.buf((long long)B * hw * hw * 4, ch);
      if (!b.dry) {
        const int hh = hw, cc = ch;
        bf16* src = cur;
        b.ops->push_back([=](const RunCtx& rc) -> int {
          OP_CUDA(launch_upsample_nearest2x(src, big, B, hh, hh, cc, rc.stream));
          return 0;
        });
      }
      b.rel(cur);
      hw *= 2;
      int npad = 0;
      const bf16* w = b.conv_w(up + ".upsamplers.0.conv.weight", &npad);
      bf16* o = b.buf((long long)B * hw * hw, ch);
      Epilogue e;
      e.bias = b.f32_pad(up + ".upsamplers.0.conv.bias", npad);
      e.n_out = ch;
      e.out = o;
      e.ld_out = ch;
      cur_sums = (npad == ch && b.gn_fusable(ch, G, (long long)hw * hw)) ? b.gn_slot(B, G) : nullptr;
      Builder::want_gn_stats(e, cur_sums, ch, G, (long long)hw * hw);
      b.conv3(big, B, hw, hw, ch, w, npad, 1, 1, e);
      b.rel(big);
      cur = o;
    }
  }
  const long long M = (long long)B * hw * hw;
  bf16* t = b.buf(M, ch);
  if (cur_sums) b.groupnorm_from_sums(cur, t, V + "conv_norm_out", B, hw * hw, ch, G, eps, true, cur_sums);
  else b.groupnorm(cur, t, V + "conv_norm_out", B, hw * hw, ch, G, eps, true);
  b.rel(cur);
  {
    int npad = 0;
    const bf16* w = b.conv_w(V + "conv_out.weight", &npad);
    const float* bias = b.f32_pad(V + "conv_out.bias", npad);
    if (!b.dry && !b.err) {
      // the destination comes with the call (rc.dec_image_out): the launch is built there, once per destination
      const int S = hw, cc = ch;
      struct Cached { float* dst = nullptr; GemmLaunch g; };
      auto cache = std::make_shared<Cached>();
      b.ops->tag(kKindGemm, 2.0 * (double)M * 3.0 * 9.0 * ch, "vae decoder conv_out -> fp32 NHWC image");
      b.ops->push_back([=](const RunCtx& rc) -> int {
        if (cache->dst != rc.dec_image_out) {
          Epilogue e;
          e.bias = bias;
          e.n_out = 3;
          e.out_f32 = rc.dec_image_out;
          e.ld_out_f32 = 3;
          GDF_TRY(build_conv3x3(&cache->g, t, B, S, S, cc, w, npad, 1, 1, e));
          cache->dst = rc.dec_image_out;
        }
        OP_CUDA(launch_gemm(cache->g, rc.stream));
        return 0;
      });
    }
  }
  b.rel(t);
  return b.err;
}

static int build_vae(Builder& b) {
  gdf_handle_s* h = b.h;
  const gdf_vae_arch& a = h->va;
  const int B = h->B, G = a.norm_num_groups;
  const float eps = a.norm_eps;
  const std::string V = "vae.encoder.";
  b.ops = &h->vae_ops;
  b.ops->scope = "vae-encoder";
  int hw = h->img;
  int ch = a.block_out_channels[0];
  b.gn_begin(B, G, 64);
  // statistics of `cur` for the GroupNorm that reads it next (null: that GroupNorm runs its own statistics pass)
  float* cur_sums = b.gn_fusable(ch, G, (long long)hw * hw) ? b.gn_slot(B, G) : nullptr;
  // conv_in: image fp32 NCHW -> NHWC bf16. Fused kernel (conv_in_sm100.cu: the 27-tap operand is built in shared memory)
  // when the shape allows; otherwise im2col (K = 27 -> 64) + GEMM. GDF_CONV_IN_FUSED=0 forces the latter (A/B timing).
  bf16* cur = b.buf((long long)B * hw * hw, ch);
  bool fused_in = conv_in_fused_supported(a.in_channels, ch, hw);
  {
    const char* ev = getenv("GDF_CONV_IN_FUSED");
    if (ev && ev[0] == '0') fused_in = false;
    if (cur_sums && !((ch / G) == 4 || (ch / G) == 8 || (ch / G) == 16)) fused_in = false;
  }
  if (fused_in) {
    int npad = 0;
    const bf16* w = b.conv_w(V + "conv_in.weight", &npad, 64);
    const float* bias = b.f32_pad(V + "conv_in.bias", npad);
    if (!b.err && !b.check_mat(w, ch, 64, "VAE conv_in weight")) return b.err;
    if (!b.dry && !b.err) {
      const int S = hw, cpg = ch / G;
      float* sums = cur_sums;
      b.ops->tag(kKindGemm, 2.0 * B * S * S * (double)ch * 27.0, "conv_in fused (image -> NHWC, K=27) N=" + std::to_string(ch));
      b.ops->push_back([=](const RunCtx& rc) -> int {
        return launch_conv_in_fused(rc.images, w, bias, cur, B, S, S, ch, sums, cpg, G, rc.stream);
      });
    }
  } else {
    bf16* col = b.buf((long long)B * hw * hw, 64);
    if (!b.dry) {
      const int S = hw, cin = a.in_channels;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        OP_CUDA(launch_im2col_small(rc.images, nullptr, col, B, S, S, cin, rc.stream));
        return 0;
      });
    }
    int npad = 0;
    const bf16* w = b.conv_w(V + "conv_in.weight", &npad, 64);
    Epilogue e;
    e.bias = b.f32_pad(V + "conv_in.bias", npad);
    e.n_out = ch;
    e.out = cur;
    e.ld_out = ch;
    Builder::want_gn_stats(e, cur_sums, ch, G, (long long)hw * hw);
    b.linear(col, (long long)B * hw * hw, 64, 64, w, npad, e);
    b.rel(col);
  }
  for (int i = 0; i < a.num_levels; ++i) {
    const int cout = a.block_out_channels[i];
    for (int j = 0; j < a.layers_per_block; ++j) {
      bf16* o = b.buf((long long)B * hw * hw, cout);
      Dest d;
      d.out = o;
      d.ld = cout;
      // the output feeds a GroupNorm (next resnet / mid block) unless a downsampling conv follows
      const bool last_of_level = (j == a.layers_per_block - 1) && (i != a.num_levels - 1);
      d.gn_sums = (!last_of_level && b.gn_fusable(cout, G, (long long)hw * hw)) ? b.gn_slot(B, G) : nullptr;
      emit_resnet(b, V + "down_blocks." + std::to_string(i) + ".resnets." + std::to_string(j), "", cur, B, hw, hw, ch,
                  cout, nullptr, 0, G, eps, d, cur_sums);
      b.rel(cur);
      cur = o;
      cur_sums = d.gn_sums;
      ch = cout;
    }
    if (i != a.num_levels - 1) {  // Downsample2D(padding=0): F.pad (0,1,0,1) + conv s2 (downsampling.py:141-147)
      const std::string wp = V + "down_blocks." + std::to_string(i) + ".downsamplers.0.conv";
      int npad = 0;
      const bf16* w = b.conv_w(wp + ".weight", &npad);
      const int ho = hw / 2;
      bf16* o = b.buf((long long)B * ho * ho, ch);
      Epilogue e;
      e.bias = b.f32_pad(wp + ".bias", npad);
      e.n_out = ch;
      e.out = o;
      e.ld_out = ch;
      cur_sums = (npad == ch && b.gn_fusable(ch, G, (long long)ho * ho)) ? b.gn_slot(B, G) : nullptr;
      Builder::want_gn_stats(e, cur_sums, ch, G, (long long)ho * ho);
      b.conv3(cur, B, hw, hw, ch, w, npad, 2, 0, e);
      b.rel(cur);
      cur = o;
      hw = ho;
    }
  }
  const long long M = (long long)B * hw * hw;
  const int N = hw * hw;
  emit_vae_mid_block(b, "vae.encoder.mid_block", cur, cur_sums, B, hw, ch, G, eps);
  // conv_norm_out + SiLU + (conv_out . quant_conv folded into one 3x3 conv) -> fp32 moments [M, 8]
  bf16* t = b.buf(M, ch);
  if (cur_sums) b.groupnorm_from_sums(cur, t, V + "conv_norm_out", B, N, ch, G, eps, true, cur_sums);
  else b.groupnorm(cur, t, V + "conv_norm_out", B, N, ch, G, eps, true);
  b.rel(cur);
  const int nm = 2 * a.latent_channels;
  {
    const std::string key = "vae#folded_conv_out";
    if (!h->raw.count(key + ".weight")) {
      const RawW* w = b.raw(V + "conv_out.weight");
      const RawW* bi = b.raw(V + "conv_out.bias");
      const RawW* wq = b.raw("vae.quant_conv.weight");
      const RawW* bqv = b.raw("vae.quant_conv.bias");
      if (w && bi && wq && bqv) {
        std::vector<float> hw_(w->numel), hb(bi->numel), hq(wq->numel), hbq(bqv->numel);
        cudaMemcpy(hw_.data(), w->ptr, w->numel * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hb.data(), bi->ptr, bi->numel * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hq.data(), wq->ptr, wq->numel * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hbq.data(), bqv->ptr, bqv->numel * 4, cudaMemcpyDeviceToHost);
        const int64_t per = w->numel / nm;
        std::vector<float> fw(w->numel, 0.f), fb(nm, 0.f);
        for (int o = 0; o < nm; ++o) {
          double bacc = hbq[o];
          for (int m = 0; m < nm; ++m) {
            const float q = hq[o * nm + m];
            bacc += (double)q * hb[m];
            for (int64_t k = 0; k < per; ++k) fw[o * per + k] += q * hw_[m * per + k];
          }
          fb[o] = (float)bacc;
        }
        RawW fwr, fbr;
        fwr.shape = w->shape;
        fwr.numel = w->numel;
        fwr.ptr = static_cast<float*>(b.dev_alloc(w->numel * 4));
        fbr.shape = {nm};
        fbr.numel = nm;
        fbr.ptr = static_cast<float*>(b.dev_alloc(nm * 4));
        if (fwr.ptr && fbr.ptr) {
          cudaMemcpy(fwr.ptr, fw.data(), w->numel * 4, cudaMemcpyHostToDevice);
          cudaMemcpy(fbr.ptr, fb.data(), nm * 4, cudaMemcpyHostToDevice);
          h->raw[key + ".weight"] = fwr;
          h->raw[key + ".bias"] = fbr;
        }
      }
    }
    int npad = 0;
    const bf16* w = b.conv_w(key + ".weight", &npad);
    float* moments = b.fbuf(M * nm);
    Epilogue e;
    e.bias = b.f32_pad(key + ".bias", npad);
    e.n_out = nm;
    e.out_f32 = moments;
    e.ld_out_f32 = nm;
    b.conv3(t, B, hw, hw, ch, w, npad, 1, 1, e);
    b.rel(t);
    if (!b.dry) {
      gdf_handle_s* hh = h;
      const float sf = a.scaling_factor, shf = a.shift_factor;
      const int lc = a.latent_channels;
      b.ops->push_back([=](const RunCtx& rc) -> int {
        __half* cap = (hh->unet_in_cap >= 0 && rc.arena) ? reinterpret_cast<__half*>(rc.arena + hh->unet_in_cap)
                                                          : nullptr;
        OP_CUDA(launch_qsample(moments, rc.eps_vae, rc.eps_q, sf, shf, rc.qa, rc.qb, rc.qs, hh->latent_nhwc, cap,
                               rc.latents_out, B, N, lc, rc.stream));
        return 0;
      });
    }
  }
  return b.err;
}

static void free_plan(gdf_handle_s* h) {
  h->vae_ops.clear();
  h->unet_ops.clear();
  h->dec_ops.clear();
  h->pool.clear();
  h->sites.clear();
  h->slots.clear();
  h->requested.clear();
  h->arena_bytes = 0;
  h->planned = false;
  h->ctrl_down.clear();
  h->ctrl_mid = nullptr;
  h->n_skips = 0;
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(h->t_dev); fr(h->ctx_bf16); fr(h->add_in); fr(h->latent_nhwc); fr(h->key_bias); fr(h->g_dev);
  fr(h->sk_ws); fr(h->sk_cnt);
  h->sk_ws = nullptr; h->sk_cnt = nullptr;
  h->g_dev = nullptr;
  h->t_dev = nullptr; h->ctx_bf16 = nullptr; h->add_in = nullptr; h->latent_nhwc = nullptr; h->key_bias = nullptr;
  h->has_key_bias = false;
}

// GDF_NVTX=1: one NVTX range per block of the network (resnet / transformer block / VAE stage, named by the feature-id
// prefix the reference uses for that block) around the launches of its ops, and one per op (its label), so that an
// Nsight Systems / ncu --nvtx timeline reads in the reference's vocabulary. Off by default: zero cost in the replay loop.
static bool nvtx_enabled() {
  static const bool on = [] { const char* e = getenv("GDF_NVTX"); return e && e[0] == '1'; }();
  return on;
}
static int run_ops_nvtx(OpList& ops, const RunCtx& rc) {
  const std::string* open_scope = nullptr;
  int r = GDF_OK;
  for (size_t i = 0; i < ops.size() && r == GDF_OK; ++i) {
    if (!open_scope || *open_scope != ops.scopes[i]) {
      if (open_scope) nvtxRangePop();
      nvtxRangePushA(ops.scopes[i].empty() ? "gdf" : ops.scopes[i].c_str());
      open_scope = &ops.scopes[i];
    }
    nvtxRangePushA(ops.labels[i].empty() ? "op" : ops.labels[i].c_str());
    r = ops.fns[i](rc);
    nvtxRangePop();
  }
  if (open_scope) nvtxRangePop();
  return r;
}

static int run_ops(gdf_handle_s* h, OpList& ops, const RunCtx& rc) {
  if (!h->profile) {
    if (nvtx_enabled()) return run_ops_nvtx(ops, rc);
    for (auto& op : ops.fns) GDF_TRY(op(rc));
    return GDF_OK;
  }
  // profiling pass: one CUDA-event pair per op on the launching stream (ops are serialised on it)
  const size_t n = ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  cudaEventRecord(ev[0], rc.stream);
  int r = GDF_OK;
  for (size_t i = 0; i < n && r == GDF_OK; ++i) {
    r = ops.fns[i](rc);
    cudaEventRecord(ev[i + 1], rc.stream);
  }
  cudaStreamSynchronize(rc.stream);
  if (r == GDF_OK) {
    for (size_t i = 0; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      ops.last_ms[i] = ms;
      h->prof_ms[ops.kinds[i]] += ms;
      h->prof_flops[ops.kinds[i]] += ops.flops[i];
      h->prof_launches[ops.kinds[i]] += 1;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return r;
}

}  // namespace gdf

// ============================================================================================ C ABI
extern "C" {

int gdf_create(const gdf_unet_arch* unet, const gdf_vae_arch* vae, int device, gdf_handle* out) {
  if (!unet || !vae || !out) return fail(GDF_ERR_INVALID, "gdf_create: null argument");
  if (unet->num_levels < 2 || unet->num_levels > GDF_MAX_LEVELS || vae->num_levels > GDF_MAX_LEVELS)
    return fail(GDF_ERR_INVALID, "gdf_create: num_levels out of range");
  GDF_CUDA(cudaSetDevice(device));
  gdf_handle_s* h = new gdf_handle_s();
  h->ua = *unet;
  h->va = *vae;
  h->device = device;
  *out = h;
  return GDF_OK;
}

int gdf_create_dit(const gdf_dit_arch* dit, const gdf_vae_arch* vae, int device, gdf_handle* out) {
  if (!dit || !vae || !out) return fail(GDF_ERR_INVALID, "gdf_create_dit: null argument");
  if (dit->num_layers < 1 || dit->num_heads < 1 || dit->patch_size < 1 || vae->num_levels > GDF_MAX_LEVELS)
    return fail(GDF_ERR_INVALID, "gdf_create_dit: bad architecture");
  GDF_CUDA(cudaSetDevice(device));
  gdf_handle_s* h = new gdf_handle_s();
  memset(&h->ua, 0, sizeof(h->ua));
  h->da = *dit;
  h->is_dit = true;
  h->ua.in_channels = dit->in_channels;              // shared latent plumbing (gdf_encode_*, q_sample)
  h->ua.cross_attention_dim = dit->caption_channels; // width of the fp32 -> bf16 context staging buffer
  h->va = *vae;
  h->device = device;
  h->ctx_len = 300;
  *out = h;
  return GDF_OK;
}

int gdf_create_flux(const gdf_flux_arch* flux, const gdf_vae_arch* vae, int device, gdf_handle* out) {
  if (!flux || !vae || !out) return fail(GDF_ERR_INVALID, "gdf_create_flux: null argument");
  if (flux->num_layers < 0 || flux->num_single_layers < 0 || flux->num_heads < 1 || flux->in_channels < 4 ||
      vae->num_levels > GDF_MAX_LEVELS)
    return fail(GDF_ERR_INVALID, "gdf_create_flux: bad architecture");
  GDF_CUDA(cudaSetDevice(device));
  gdf_handle_s* h = new gdf_handle_s();
  memset(&h->ua, 0, sizeof(h->ua));
  memset(&h->da, 0, sizeof(h->da));
  h->fa = *flux;
  h->is_flux = true;
  h->ua.in_channels = flux->in_channels / 4;             // latent channels (shared latent plumbing, q_sample)
  h->ua.cross_attention_dim = flux->joint_attention_dim; // width of the fp32 -> bf16 context staging buffer
  h->va = *vae;
  h->device = device;
  h->ctx_len = 512;
  *out = h;
  return GDF_OK;
}

int gdf_destroy(gdf_handle h) {
  if (!h) return GDF_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_plan(h);
  for (auto& kv : h->raw)
    if (kv.second.ptr && kv.first.find('#') == std::string::npos) cudaFree(kv.second.ptr);
  for (void* p : h->owned) cudaFree(p);
  delete h;
  return GDF_OK;
}

int gdf_load_weights(gdf_handle h, const char* const* names, const void* const* ptrs_dev, const int64_t* shapes,
                     const int* ranks, int n, void* stream) {
  if (!h) return fail(GDF_ERR_INVALID, "null handle");
  if (n > 0 && (!names || !ptrs_dev || !shapes || !ranks)) return fail(GDF_ERR_INVALID, "gdf_load_weights: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GDF_CUDA(cudaSetDevice(h->device));
  // Everything derived from the weights held so far is dropped: the plan (its op list holds raw pointers to biases and
  // norm affines), every packed bf16 / folded tensor (the cache is keyed by name and would silently keep the OLD
  // values) and the shape registry. The caller re-finalises and re-plans.
  if (h->planned || !h->packed.empty()) {
    GDF_CUDA(cudaDeviceSynchronize());
    free_plan(h);
    for (void* p : h->owned) cudaFree(p);
    h->owned.clear();
    h->packed.clear();
    h->mat_dims.clear();
    h->vec_len.clear();
    for (auto it = h->raw.begin(); it != h->raw.end();) {   // derived entries ("vae#folded_conv_out.*") lived in `owned`
      if (it->first.find('#') != std::string::npos) it = h->raw.erase(it);
      else ++it;
    }
  }
  const int64_t* sp = shapes;
  for (int i = 0; i < n; ++i) {
    if (ranks[i] < 0 || ranks[i] > 8) return fail(GDF_ERR_SHAPE, "gdf_load_weights: '%s' has rank %d", names[i], ranks[i]);
    RawW w;
    w.numel = 1;
    for (int d = 0; d < ranks[i]; ++d) {
      w.shape.push_back(sp[d]);
      w.numel *= sp[d];
    }
    sp += ranks[i];
    GDF_CUDA(cudaMalloc(&w.ptr, (size_t)(w.numel ? w.numel : 1) * 4));
    GDF_CUDA(cudaMemcpyAsync(w.ptr, ptrs_dev[i], (size_t)w.numel * 4, cudaMemcpyDeviceToDevice, st));
    auto it = h->raw.find(names[i]);
    if (it != h->raw.end()) {
      if (it->second.ptr) cudaFree(it->second.ptr);
      h->raw.erase(it);
    }
    h->raw[names[i]] = w;
  }
  GDF_CUDA(cudaStreamSynchronize(st));
  h->finalized = false;
  return GDF_OK;
}

int gdf_finalize_weights(gdf_handle h, void* stream) {
  (void)stream;
  if (!h) return fail(GDF_ERR_INVALID, "null handle");
  // dry walk: touches (and packs) every weight the architecture needs; reports the first missing name
  const int B0 = h->B, img0 = h->img, L0 = h->L;
  h->B = 1;
  h->img = 64;
  h->L = 8;
  Builder b(h, true);
  std::vector<Site> keep_sites = h->sites;
  int r = build_vae(b);
  // optional: only `vae-out` decodes (UNet families: 4 latent channels; a 16-channel Flux decoder is never walked)
  if (!r && h->raw.count("vae.decoder.conv_in.weight") && 9 * h->va.latent_channels <= 64) r = build_vae_decoder(b);
  if (!r) r = h->is_flux ? build_flux(b) : h->is_dit ? build_dit(b) : build_unet(b);
  h->sites = keep_sites;
  h->B = B0;
  h->img = img0;
  h->L = L0;
  GDF_CUDA(cudaDeviceSynchronize());
  if (r) return r;
  // The fp32 originals of the matrices are only needed for packing: release them (SDXL: ~10 GB, Flux: ~48 GB). What
  // the op lists read in place (biases, norm affines, conditioning MLPs, position tables) was marked `keep` by the walk.
  const char* keep_all = getenv("GDF_KEEP_FP32_WEIGHTS");
  if (!(keep_all && keep_all[0] == '1')) {
    for (auto& kv : h->raw) {
      RawW& w = kv.second;
      if (w.keep || !w.packed_from || !w.ptr || w.shape.size() < 2 || kv.first.find('#') != std::string::npos)
        continue;
      cudaFree(w.ptr);
      w.ptr = nullptr;
    }
  }
  h->finalized = true;
  return GDF_OK;
}

int gdf_plan(gdf_handle h, const char* const* feature_ids, int n_ids, int batch, int img_size, gdf_slot* slots_out,
             int64_t* arena_bytes_out) {
  if (!h) return fail(GDF_ERR_INVALID, "null handle");
  if (!h->finalized) return fail(GDF_ERR_INVALID, "gdf_plan: call gdf_finalize_weights first");
  if (batch < 1 || img_size < 64 || img_size % 64 != 0)
    return fail(GDF_ERR_SHAPE, "gdf_plan: batch %d / img_size %d unsupported (img_size must be a multiple of 64)",
                batch, img_size);
  GDF_CUDA(cudaSetDevice(h->device));
  GDF_CUDA(cudaDeviceSynchronize());
  free_plan(h);
  h->B = batch;
  h->img = img_size;
  h->L = img_size / 8;
  h->slots.assign(n_ids, gdf_slot{-1, 0, 0, 0, -1});
  for (int i = 0; i < n_ids; ++i) {
    const std::string id = feature_ids[i];
    if (id == "vae-out" || id == "attn")
      return fail(GDF_ERR_UNSUPPORTED, "feature id '%s': vae-out needs the VAE decoder / `attn` is assembled by the host "
                  "from the #attnmean slots", id.c_str());
    h->requested[id] = i;
  }
  const gdf_unet_arch& a = h->ua;
  GDF_CUDA(cudaMalloc(&h->t_dev, (size_t)batch * 4));
  GDF_CUDA(cudaMalloc(&h->ctx_bf16, (size_t)batch * h->ctx_len * a.cross_attention_dim * 2));
  GDF_CUDA(cudaMalloc(&h->latent_nhwc, (size_t)batch * h->L * h->L * a.in_channels * 2));
  if (a.addition_time_embed_dim > 0)
    GDF_CUDA(cudaMalloc(&h->add_in, (size_t)batch * a.projection_class_embeddings_input_dim * 4));
  Builder b(h, false);
  b.gn_ws = static_cast<float*>(h->pool.acquire(gn_workspace_floats(batch, 64) * 4));
  // UNet first: registers the capture sites (incl. unet-in, written by the q_sample kernel of the VAE pass)
  if (h->is_dit) GDF_CUDA(cudaMalloc(&h->key_bias, (size_t)batch * h->ctx_len * 4));
  if (h->is_flux) GDF_CUDA(cudaMalloc(&h->g_dev, (size_t)batch * 4));
  int r = h->is_flux ? build_flux(b) : h->is_dit ? build_dit(b) : build_unet(b);
  if (!r) r = build_vae(b);
  if (r) {
    free_plan(h);
    return r;
  }
  // every requested id must have been produced by the walk (cross-k/v are accepted and never stored)
  for (auto& kv : h->requested) {
    if (h->slots[kv.second].offset_bytes >= 0) continue;
    const std::string& id = kv.first;
    if (id.find("cross-k") != std::string::npos || id.find("cross-v") != std::string::npos) continue;
    std::string msg = "unknown feature id '" + id + "' for this architecture";
    free_plan(h);
    return fail(GDF_ERR_INVALID, "%s", msg.c_str());
  }
  for (int i = 0; i < n_ids; ++i) {
    // duplicates in the caller's list share the first slot
    slots_out[i] = h->slots[h->requested[feature_ids[i]]];
  }
  if (arena_bytes_out) *arena_bytes_out = h->arena_bytes > 0 ? h->arena_bytes : 256;
  h->planned = true;
  ++h->plan_generation;
  h->gpu_launches = (int)(h->vae_ops.size() + h->unet_ops.size());
  return GDF_OK;
}

int gdf_plan_decoder(gdf_handle h) {
  if (!h || !h->planned) return fail(GDF_ERR_INVALID, "gdf_plan_decoder: call gdf_plan first");
  if (!h->dec_ops.fns.empty()) return GDF_OK;
  if (!h->raw.count("vae.decoder.conv_in.weight"))
    return fail(GDF_ERR_MISSING_WEIGHT, "gdf_plan_decoder: no VAE decoder weights ('vae.decoder.*') were loaded");
  GDF_CUDA(cudaSetDevice(h->device));
  Builder b(h, false);
  b.gn_ws = static_cast<float*>(h->pool.acquire(gn_workspace_floats(h->B, 64) * 4));
  const int r = build_vae_decoder(b);
  if (r) {
    h->dec_ops.clear();
    return r;
  }
  return GDF_OK;
}

int gdf_decode_latents(gdf_handle h, const void* latents_dev, float c_latent, const void* model_out_dev, float c_model,
                       void* image_out_dev, void* stream) {
  if (!h || !h->planned || h->dec_ops.fns.empty()) return fail(GDF_ERR_INVALID, "gdf_decode_latents: call gdf_plan_decoder first");
  if (!latents_dev || !image_out_dev) return fail(GDF_ERR_INVALID, "gdf_decode_latents: null argument");
  RunCtx rc;
  rc.stream = static_cast<cudaStream_t>(stream);
  rc.dec_latents = static_cast<const float*>(latents_dev);
  rc.dec_model_out = static_cast<const float*>(model_out_dev);
  rc.dec_c_latent = c_latent;
  rc.dec_c_model = c_model;
  rc.dec_image_out = static_cast<float*>(image_out_dev);
  GDF_TRY(run_ops(h, h->dec_ops, rc));
  return GDF_OK;
}

uint64_t gdf_plan_generation(gdf_handle h) { return (h && h->planned) ? h->plan_generation : 0; }

int gdf_control_residual_shapes(gdf_handle h, int* channels_out, int* sides_out, int max_n) {
  if (!h || !h->planned || h->is_dit || h->is_flux) return fail(GDF_ERR_INVALID, "gdf_control_residual_shapes: no UNet plan");
  for (int i = 0; i < h->n_skips && i < max_n; ++i) {
    if (channels_out) channels_out[i] = h->skip_shapes[i].first;
    if (sides_out) sides_out[i] = h->skip_shapes[i].second;
  }
  return h->n_skips;
}

int gdf_set_control_residuals(gdf_handle h, const void* const* down_dev, int n_down, const void* mid_dev) {
  if (!h || !h->planned || h->is_dit || h->is_flux) return fail(GDF_ERR_INVALID, "gdf_set_control_residuals: no UNet plan");
  if (n_down != 0 && n_down != h->n_skips)
    return fail(GDF_ERR_SHAPE, "gdf_set_control_residuals: %d down residuals, this UNet has %d skip tensors", n_down,
                h->n_skips);
  h->ctrl_down.clear();
  for (int i = 0; i < n_down; ++i) h->ctrl_down.push_back(static_cast<const float*>(down_dev[i]));
  h->ctrl_mid = static_cast<const float*>(mid_dev);
  return GDF_OK;
}

int gdf_encode_noise(gdf_handle h, const void* images_dev, const void* eps_vae_dev, const void* eps_q_dev,
                     float sqrt_alpha_bar, float sqrt_one_minus_alpha_bar, float input_scale, void* latents_out_dev,
                     void* stream) {
  if (!h || !h->planned) return fail(GDF_ERR_INVALID, "gdf_encode_noise: no plan");
  if (!images_dev || !eps_vae_dev || !eps_q_dev) return fail(GDF_ERR_INVALID, "gdf_encode_noise: null input");
  RunCtx rc;
  rc.stream = static_cast<cudaStream_t>(stream);
  rc.images = static_cast<const float*>(images_dev);
  rc.eps_vae = static_cast<const float*>(eps_vae_dev);
  rc.eps_q = static_cast<const float*>(eps_q_dev);
  rc.qa = sqrt_alpha_bar;
  rc.qb = sqrt_one_minus_alpha_bar;
  rc.qs = input_scale;
  rc.latents_out = static_cast<float*>(latents_out_dev);
  rc.arena = nullptr;  // unet-in is re-captured by gdf_denoise_capture from the stored latent
  GDF_TRY(run_ops(h, h->vae_ops, rc));
  return GDF_OK;
}

int gdf_encode_latents(gdf_handle h, const void* latents_dev, const void* eps_q_dev, float sqrt_alpha_bar,
                       float sqrt_one_minus_alpha_bar, float input_scale, void* latents_out_dev, void* stream) {
  if (!h || !h->planned) return fail(GDF_ERR_INVALID, "gdf_encode_latents: no plan");
  if (!latents_dev) return fail(GDF_ERR_INVALID, "gdf_encode_latents: null input");
  const int HW = h->L * h->L, C = h->ua.in_channels;
  latents_qsample_kernel<<<256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(latents_dev), static_cast<const float*>(eps_q_dev), sqrt_alpha_bar,
      sqrt_one_minus_alpha_bar, input_scale, h->latent_nhwc, static_cast<float*>(latents_out_dev), h->B, HW, C);
  GDF_CUDA(cudaGetLastError());
  return GDF_OK;
}

static int check_arena(gdf_handle h, int64_t arena_bytes, const char* who) {
  if (arena_bytes < h->arena_bytes)
    return fail(GDF_ERR_SHAPE, "%s: the arena holds %lld bytes, the current plan (generation %llu) writes %lld", who,
                (long long)arena_bytes, (unsigned long long)h->plan_generation, (long long)h->arena_bytes);
  return GDF_OK;
}

int gdf_denoise_capture(gdf_handle h, float timestep, const void* ctx_dev, int ctx_len, const void* pooled_dev,
                        const void* add_time_ids_dev, void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev,
                        void* stream) {
  if (!h || !h->planned) return fail(GDF_ERR_INVALID, "gdf_denoise_capture: no plan");
  GDF_TRY(check_arena(h, arena_bytes, "gdf_denoise_capture"));
  if (ctx_len != h->ctx_len)
    return fail(GDF_ERR_SHAPE, "gdf_denoise_capture: ctx_len %d, plan was built for %d", ctx_len, h->ctx_len);
  if (h->is_dit) return fail(GDF_ERR_INVALID, "gdf_denoise_capture: this handle holds a DiT, use gdf_denoise_capture_dit");
  if (h->is_flux) return fail(GDF_ERR_INVALID, "gdf_denoise_capture: this handle holds Flux, use gdf_denoise_capture_flux");
  if (!ctx_dev || !arena_dev) return fail(GDF_ERR_INVALID, "gdf_denoise_capture: null input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RunCtx rc;
  rc.stream = st;
  rc.arena = static_cast<char*>(arena_dev);
  rc.noise_pred_out = static_cast<float*>(noise_pred_out_dev);
  h->pooled = static_cast<const float*>(pooled_dev);
  h->time_ids_dev = const_cast<float*>(static_cast<const float*>(add_time_ids_dev));
  fill_f32_kernel<<<(h->B + 255) / 256, 256, 0, st>>>(h->t_dev, timestep, h->B);
  GDF_CUDA(cudaGetLastError());
  GDF_CUDA(launch_cast_f32_to_bf16(static_cast<const float*>(ctx_dev), h->ctx_bf16,
                                   (long long)h->B * h->ctx_len * h->ua.cross_attention_dim, st));
  if (h->unet_in_cap >= 0)  // unet-in (unet_2d_condition.py:1169-1170): the scaled latent, fp16 token-major
    GDF_CUDA(launch_cast_bf16_to_f16(h->latent_nhwc, reinterpret_cast<__half*>(rc.arena + h->unet_in_cap),
                                     (long long)h->B * h->L * h->L * h->ua.in_channels, st));
  GDF_TRY(run_ops(h, h->unet_ops, rc));
  return GDF_OK;
}

int gdf_denoise_capture_dit(gdf_handle h, float timestep, const void* ctx_dev, int ctx_len, const void* ctx_mask_dev,
                            void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev, void* stream) {
  if (!h || !h->planned || !h->is_dit) return fail(GDF_ERR_INVALID, "gdf_denoise_capture_dit: no DiT plan");
  GDF_TRY(check_arena(h, arena_bytes, "gdf_denoise_capture_dit"));
  if (ctx_len != h->ctx_len)
    return fail(GDF_ERR_SHAPE, "gdf_denoise_capture_dit: ctx_len %d, plan was built for %d", ctx_len, h->ctx_len);
  if (!ctx_dev || !arena_dev) return fail(GDF_ERR_INVALID, "gdf_denoise_capture_dit: null input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RunCtx rc;
  rc.stream = st;
  rc.arena = static_cast<char*>(arena_dev);
  rc.noise_pred_out = static_cast<float*>(noise_pred_out_dev);
  fill_f32_kernel<<<(h->B + 255) / 256, 256, 0, st>>>(h->t_dev, timestep, h->B);
  GDF_CUDA(cudaGetLastError());
  GDF_CUDA(launch_cast_f32_to_bf16(static_cast<const float*>(ctx_dev), h->ctx_bf16,
                                   (long long)h->B * h->ctx_len * h->da.caption_channels, st));
  h->has_key_bias = ctx_mask_dev != nullptr;
  if (ctx_mask_dev)
    GDF_CUDA(launch_mask_to_bias(static_cast<const float*>(ctx_mask_dev), h->key_bias, h->B * h->ctx_len, st));
  GDF_TRY(run_ops(h, h->unet_ops, rc));
  return GDF_OK;
}

int gdf_denoise_capture_flux(gdf_handle h, float sigma, float guidance, const void* ctx_dev, int ctx_len,
                             const void* pooled_dev, const void* rope_cos_dev, const void* rope_sin_dev,
                             void* arena_dev, int64_t arena_bytes, void* noise_pred_out_dev, void* stream) {
  if (!h || !h->planned || !h->is_flux) return fail(GDF_ERR_INVALID, "gdf_denoise_capture_flux: no Flux plan");
  GDF_TRY(check_arena(h, arena_bytes, "gdf_denoise_capture_flux"));
  if (ctx_len != h->ctx_len)
    return fail(GDF_ERR_SHAPE, "gdf_denoise_capture_flux: ctx_len %d, plan was built for %d", ctx_len, h->ctx_len);
  if (!ctx_dev || !pooled_dev || !rope_cos_dev || !rope_sin_dev || !arena_dev)
    return fail(GDF_ERR_INVALID, "gdf_denoise_capture_flux: null input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RunCtx rc;
  rc.stream = st;
  rc.arena = static_cast<char*>(arena_dev);
  rc.noise_pred_out = static_cast<float*>(noise_pred_out_dev);
  h->pooled = static_cast<const float*>(pooled_dev);
  h->rope_cos = static_cast<const float*>(rope_cos_dev);
  h->rope_sin = static_cast<const float*>(rope_sin_dev);
  // transformer_flux.py:455-459: timestep = timestep * 1000, guidance = guidance * 1000
  fill_f32_kernel<<<(h->B + 255) / 256, 256, 0, st>>>(h->t_dev, sigma * 1000.f, h->B);
  fill_f32_kernel<<<(h->B + 255) / 256, 256, 0, st>>>(h->g_dev, guidance * 1000.f, h->B);
  GDF_CUDA(cudaGetLastError());
  GDF_CUDA(launch_cast_f32_to_bf16(static_cast<const float*>(ctx_dev), h->ctx_bf16,
                                   (long long)h->B * h->ctx_len * h->fa.joint_attention_dim, st));
  GDF_TRY(run_ops(h, h->unet_ops, rc));
  return GDF_OK;
}

/* Profiling pass: while enabled, encode/denoise record a CUDA-event pair around every kernel launch and
 * accumulate device time / algorithmic FLOPs / launch counts per kernel kind (0 tcgen05 GEMM+conv, 1 attention,
 * 2 GroupNorm, 3 LayerNorm, 4 other). Enabling resets the counters. */
int gdf_profile(gdf_handle h, int enable) {
  if (!h) return fail(GDF_ERR_INVALID, "null handle");
  h->profile = enable != 0;
  if (enable)
    for (int i = 0; i < kNumKinds; ++i) {
      h->prof_ms[i] = 0.f;
      h->prof_flops[i] = 0.0;
      h->prof_launches[i] = 0;
    }
  return GDF_OK;
}
int gdf_profile_read(gdf_handle h, float* ms_out, double* flops_out, int* launches_out) {
  if (!h) return fail(GDF_ERR_INVALID, "null handle");
  for (int i = 0; i < kNumKinds; ++i) {
    ms_out[i] = h->prof_ms[i];
    flops_out[i] = h->prof_flops[i];
    launches_out[i] = h->prof_launches[i];
  }
  return GDF_OK;
}
/* Writes the per-launch table of the last profiling pass (phase, index, kind, ms, GFLOP, label) as CSV. */
int gdf_profile_dump(gdf_handle h, const char* path) {
  if (!h || !path) return fail(GDF_ERR_INVALID, "gdf_profile_dump");
  FILE* f = fopen(path, "w");
  if (!f) return fail(GDF_ERR_INVALID, "cannot open %s", path);
  fprintf(f, "phase,index,kind,ms,gflop,tflops,label\n");
  const char* names[2] = {"vae", "unet"};
  OpList* lists[2] = {&h->vae_ops, &h->unet_ops};
  for (int l = 0; l < 2; ++l)
    for (size_t i = 0; i < lists[l]->size(); ++i) {
      const float ms = lists[l]->last_ms[i];
      const double fl = lists[l]->flops[i];
      fprintf(f, "%s,%zu,%d,%.4f,%.3f,%.1f,%s\n", names[l], i, lists[l]->kinds[i], ms, fl / 1e9,
              ms > 0 ? fl / (ms * 1e-3) / 1e12 : 0.0, lists[l]->labels[i].c_str());
    }
  fclose(f);
  return GDF_OK;
}
int gdf_num_launches(gdf_handle h) { return h ? h->gpu_launches : 0; }
int64_t gdf_workspace_bytes(gdf_handle h) { return h ? (int64_t)h->pool.total() : 0; }
int gdf_set_ctx_len(gdf_handle h, int ctx_len) {
  if (!h || ctx_len < 1) return fail(GDF_ERR_INVALID, "gdf_set_ctx_len");
  if (ctx_len != h->ctx_len) {
    h->ctx_len = ctx_len;
    free_plan(h);
  }
  return GDF_OK;
}

}  // extern "C"
