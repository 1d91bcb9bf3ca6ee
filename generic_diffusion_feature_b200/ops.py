"""Torch-tensor front end of the op-level C ABI (gdf_op_*). torch is only the device allocator here."""
import ctypes

import torch

from . import _lib
from ._lib import CaptureSeg, Epilogue, ResizeSrc, check, ptr, stream_ptr

ACT_NONE, ACT_GEGLU, ACT_GELU_TANH, ACT_SILU, ACT_RELU = 0, 1, 2, 3, 4


def make_epilogue(out=None, bias=None, bias_m=None, row_batch_bias=None, rows_per_batch=0, act=ACT_NONE,
                  col_scale=None, residual=None, out_scale=1.0, alpha=1.0, out2=None, out_f32=None, cap_pre=None,
                  caps=(), n_out=0, out_batch_stride=0, out_f16_from=0, ln_sums=None, ln_u=None, ln_eps=1e-5,
                  row_sums=None, gn_sums=None, gn_cpg=0, gn_groups=0, gn_rows_per_img=0, in_f16=False, res_f16=False,
                  k_split=None):
    e = Epilogue()
    e.alpha = alpha
    e.n_out = n_out
    e.bias_dev = ptr(bias)
    e.bias_m_dev = ptr(bias_m)
    e.row_batch_bias_dev = ptr(row_batch_bias)
    e.rows_per_batch = rows_per_batch
    e.act = act
    e.col_scale_dev = ptr(col_scale)
    if residual is not None:
        e.residual_dev = ptr(residual)
        e.ld_res = residual.stride(-2)
    e.out_scale = out_scale
    if out is not None:
        e.out_dev = ptr(out)
        e.ld_out = out.stride(-2)
        e.out_batch_stride = out_batch_stride
        e.out_f16_from = out_f16_from
    if out2 is not None:
        e.out2_dev = ptr(out2)
        e.ld_out2 = out2.stride(-2)
    if out_f32 is not None:
        e.out_f32_dev = ptr(out_f32)
        e.ld_out_f32 = out_f32.stride(-2)
    if cap_pre is not None:
        e.cap_pre_dev = ptr(cap_pre)
        e.ld_cap_pre = cap_pre.stride(-2)
    e.num_cap = len(caps)
    for i, (t, c0, c1) in enumerate(caps):
        e.cap[i] = CaptureSeg(ptr(t), c0, c1, t.stride(-2))
    e.ln_sums_dev = ptr(ln_sums)      # folded LayerNorm of the A rows (fp32 [M, 2] sums written by a producer's row_sums)
    e.ln_u_dev = ptr(ln_u)
    e.ln_eps = ln_eps
    e.row_sums_dev = ptr(row_sums)
    e.gn_sums_dev = ptr(gn_sums)      # fused GroupNorm statistics of the output: fp32 [images, groups, 2], ADDED to
    e.gn_cpg = gn_cpg
    e.gn_groups = gn_groups
    e.gn_rows_per_img = gn_rows_per_img
    e.in_f16 = int(in_f16)           # A and W are fp16 (feature stacks) instead of bf16
    e.res_f16 = int(res_f16)         # residual is fp16 (a captured feature map) instead of bf16
    if k_split is not None:          # (fp32 workspace, zeroed int32 counters) from k_split_workspace()
        ws, cnt = k_split
        e.k_split_ws_dev, e.k_split_ws_floats = ptr(ws), ws.numel()
        e.k_split_cnt_dev, e.k_split_cnt_len = ptr(cnt), cnt.numel()
    return e


def k_split_workspace(device="cuda"):
    """Workspace of the K-split tail wave (gdf_epilogue.k_split_*): always-sufficient sizes."""
    return (torch.empty(74 * 2 * 128 * 256, dtype=torch.float32, device=device),
            torch.zeros(74 * 2 * 8, dtype=torch.int32, device=device))


def linear(a, w, ep, batch=1, a_batch_stride=0, w_batch_stride=0, block_n=0, M=None):
    """a: bf16 [M, K] (row pitch a.stride(-2)), w: bf16 [N, K]."""
    lib = _lib.load()
    M = a.shape[-2] if M is None else M
    K = a.shape[-1]
    N = w.shape[-2]
    check(lib.gdf_op_linear(ptr(a), M, K, a.stride(-2), ptr(w), N, w.stride(-2), ctypes.byref(ep), batch,
                            a_batch_stride, w_batch_stride, block_n, stream_ptr()))


def pack_conv_weight(w_oihw, o_pad=None, k_pad=None):
    lib = _lib.load()
    O, I, kh, kw = w_oihw.shape
    o_pad = o_pad or ((O + 15) // 16) * 16
    k_pad = k_pad or kh * kw * I
    out = torch.empty(o_pad, k_pad, dtype=torch.bfloat16, device=w_oihw.device)
    check(lib.gdf_op_pack_conv_weight(ptr(w_oihw.float().contiguous()), ptr(out), O, o_pad, I, kh, kw, k_pad,
                                      stream_ptr()))
    return out


def pack_conv_weight_f16(w_oihw, o_pad=None, k_pad=None):
    """fp16 packing for convolutions over fp16 tensors (make_epilogue(in_f16=True))."""
    lib = _lib.load()
    O, I, kh, kw = w_oihw.shape
    o_pad = o_pad or ((O + 15) // 16) * 16
    k_pad = k_pad or kh * kw * I
    out = torch.empty(o_pad, k_pad, dtype=torch.float16, device=w_oihw.device)
    check(lib.gdf_op_pack_conv_weight_f16(ptr(w_oihw.float().contiguous()), ptr(out), O, o_pad, I, kh, kw, k_pad,
                                          stream_ptr()))
    return out


def conv3x3(x_nhwc, w_packed, ep, stride=1, pad_lo=1, block_n=0):
    lib = _lib.load()
    B, H, W, C = x_nhwc.shape
    assert x_nhwc.is_contiguous()
    check(lib.gdf_op_conv3x3(ptr(x_nhwc), B, H, W, C, ptr(w_packed), w_packed.shape[0], stride, pad_lo,
                             ctypes.byref(ep), block_n, stream_ptr()))


def conv_in_fused(img_nchw_f32, w_oihw, bias, gn_groups=0):
    """First VAE convolution fused from the image: (B,3,H,W) fp32 -> bf16 [B, H*W, N] (+ optional GroupNorm sums
    fp32 [B, gn_groups, 2] of the output)."""
    lib = _lib.load()
    B, C, H, W = img_nchw_f32.shape
    N = w_oihw.shape[0]
    wp = pack_conv_weight(w_oihw, k_pad=64)
    out = torch.empty(B, H * W, N, dtype=torch.bfloat16, device=img_nchw_f32.device)
    sums = torch.zeros(B, gn_groups, 2, dtype=torch.float32, device=out.device) if gn_groups else None
    check(lib.gdf_op_conv_in(ptr(img_nchw_f32.contiguous()), ptr(wp), ptr(bias.float().contiguous()), ptr(out), B, H, W, N,
                             ptr(sums), (N // gn_groups) if gn_groups else 0, gn_groups, stream_ptr()))
    return out, sums


def groupnorm(x, gamma, beta, groups, eps, silu):
    """x: bf16 [B, HW, C] contiguous."""
    lib = _lib.load()
    B, HW, C = x.shape
    y = torch.empty_like(x)
    ws = torch.empty(lib.gdf_op_groupnorm_workspace_floats(B, groups), dtype=torch.float32, device=x.device)
    check(lib.gdf_op_groupnorm(ptr(x), ptr(y), ptr(gamma), ptr(beta), B, HW, C, groups, eps, int(silu), ptr(ws),
                               stream_ptr()))
    return y


def layernorm(x, gamma, beta, eps, mod_scale=None, mod_shift=None, rows_per_batch=0):
    lib = _lib.load()
    M, C = x.shape
    y = torch.empty_like(x)
    check(lib.gdf_op_layernorm(ptr(x), ptr(y), ptr(gamma), ptr(beta), M, C, eps, ptr(mod_scale), ptr(mod_shift),
                               rows_per_batch, stream_ptr()))
    return y


def attention(q, k, v, B, heads, Nq, Nk, scale, head_dim=64, v_f16=False):
    """q: bf16 [B*Nq, >=heads*64] (pitch q.stride(0)); k, v: [B*Nk, ...]. Returns bf16 [B*Nq, heads*64].
    v_f16: v is a float16 tensor (tcgen05 kernel, Nk >= 128); otherwise bf16 (mma.sync kernel)."""
    lib = _lib.load()
    o = torch.zeros(B * Nq, heads * head_dim, dtype=torch.bfloat16, device=q.device)
    check(lib.gdf_op_attention(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o), o.stride(0), B,
                               heads, Nq, Nk, head_dim, scale, int(v_f16), stream_ptr()))
    return o


def attention_bias(q, k, v, B, heads, Nq, Nk, scale, head_dim, key_bias=None):
    """Any head_dim (multiple of 8, <= 160), bf16 q/k/v, optional additive fp32 key bias [B, Nk]."""
    lib = _lib.load()
    o = torch.zeros(B * Nq, heads * head_dim, dtype=torch.bfloat16, device=q.device)
    check(lib.gdf_op_attention_bias(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(v), v.stride(0), ptr(o),
                                    o.stride(0), B, heads, Nq, Nk, head_dim, scale, ptr(key_bias), stream_ptr()))
    return o


def softmax_rows_(s):
    lib = _lib.load()
    rows, cols = s.shape
    check(lib.gdf_op_softmax_rows(ptr(s), rows, cols, s.stride(0), stream_ptr()))
    return s


def upsample_nearest2x(x_nhwc):
    lib = _lib.load()
    B, H, W, C = x_nhwc.shape
    y = torch.empty(B, 2 * H, 2 * W, C, dtype=x_nhwc.dtype, device=x_nhwc.device)
    check(lib.gdf_op_upsample_nearest2x(ptr(x_nhwc), ptr(y), B, H, W, C, stream_ptr()))
    return y


def im2col_small(src, nchw_f32):
    lib = _lib.load()
    if nchw_f32:
        B, C, H, W = src.shape
    else:
        B, H, W, C = src.shape
    a = torch.empty(B * H * W, 64, dtype=torch.bfloat16, device=src.device)
    check(lib.gdf_op_im2col_small(ptr(src) if nchw_f32 else None, None if nchw_f32 else ptr(src), ptr(a), B, H, W, C,
                                  stream_ptr()))
    return a


def resize_concat(maps, out_hw, nhwc=True, nchw=False, with_sumsq=False):
    """maps: list of fp16 [B, h*w, C] token-major tensors (h == w). Returns dict with the requested outputs.

    Mirrors aggregation_network.py:62-66 (bilinear resize of every map + channel concat)."""
    lib = _lib.load()
    B = maps[0].shape[0]
    OH, OW = out_hw
    ctot = sum(m.shape[2] for m in maps)
    dev = maps[0].device
    srcs = (ResizeSrc * len(maps))()
    off = 0
    for i, m in enumerate(maps):
        assert m.dtype == torch.float16 and m.is_contiguous()
        side = int(round(m.shape[1] ** 0.5))
        assert side * side == m.shape[1]
        srcs[i] = ResizeSrc(ptr(m), side, side, m.shape[2], off)
        off += m.shape[2]
    out = {}
    o_nhwc = torch.empty(B, OH * OW, ctot, dtype=torch.float16, device=dev) if nhwc else None
    o_nchw = torch.empty(B, ctot, OH, OW, dtype=torch.float16, device=dev) if nchw else None
    sumsq = torch.empty(B, OH * OW, dtype=torch.float32, device=dev) if (with_sumsq and nhwc) else None
    check(lib.gdf_op_resize_concat(srcs, len(maps), B, OH, OW, ctot, ptr(o_nhwc), ptr(o_nchw), ptr(sumsq),
                                   stream_ptr()))
    out["nhwc"], out["nchw"], out["sumsq"] = o_nhwc, o_nchw, sumsq
    return out
