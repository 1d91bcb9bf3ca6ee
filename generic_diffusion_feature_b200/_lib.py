"""ctypes binding of libgdf_b200.so (C ABI declared in include/gdf.h).

There is no CPU fallback: if the shared library is missing the import of any compute entry point raises.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# GDF_LIB_PATH: A/B timing of two builds of the same library on one GPU box (tools/, not a fallback)
LIB_PATH = os.environ.get("GDF_LIB_PATH") or os.path.join(_PKG, "libgdf_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float

GDF_MAX_LEVELS = 4

# every symbol include/gdf.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "gdf_last_error", "gdf_abi_version",
    "gdf_create", "gdf_create_dit", "gdf_denoise_capture_dit", "gdf_create_flux", "gdf_denoise_capture_flux", "gdf_op_attention_bias", "gdf_destroy", "gdf_load_weights", "gdf_finalize_weights", "gdf_plan",
    "gdf_encode_noise", "gdf_encode_latents", "gdf_plan_decoder", "gdf_decode_latents", "gdf_denoise_capture", "gdf_set_ctx_len", "gdf_num_launches", "gdf_workspace_bytes",
    "gdf_plan_generation", "gdf_control_residual_shapes", "gdf_set_control_residuals",
    "gdf_profile", "gdf_profile_read", "gdf_profile_dump",
    "gdf_op_linear", "gdf_op_conv3x3", "gdf_op_pack_conv_weight", "gdf_op_pack_conv_weight_f16", "gdf_op_conv_in", "gdf_op_groupnorm_workspace_floats",
    "gdf_op_groupnorm", "gdf_op_layernorm", "gdf_op_attention", "gdf_op_softmax_rows", "gdf_debug_attention_trace",
    "gdf_op_upsample_nearest2x", "gdf_op_im2col_small", "gdf_op_qsample", "gdf_op_cast_f32_to_bf16",
    "gdf_op_resize_concat", "gdf_op_avgpool_nhwc", "gdf_correspond_workspace_floats", "gdf_correspond",
]


class CaptureSeg(ctypes.Structure):
    _fields_ = [("ptr_dev", c_void_p), ("col_begin", c_int), ("col_end", c_int), ("ld", c_int)]


class Epilogue(ctypes.Structure):
    _fields_ = [
        ("alpha", c_float), ("n_out", c_int),
        ("bias_dev", c_void_p), ("bias_m_dev", c_void_p), ("row_batch_bias_dev", c_void_p),
        ("rows_per_batch", c_int), ("act", c_int),
        ("col_scale_dev", c_void_p),
        ("residual_dev", c_void_p), ("ld_res", c_int),
        ("out_scale", c_float),
        ("out_dev", c_void_p), ("ld_out", c_int), ("out_batch_stride", c_int64), ("out_f16_from", c_int),
        ("out2_dev", c_void_p), ("ld_out2", c_int),
        ("out_f32_dev", c_void_p), ("ld_out_f32", c_int),
        ("cap_pre_dev", c_void_p), ("ld_cap_pre", c_int),
        ("cap", CaptureSeg * 3), ("num_cap", c_int),
        ("ln_sums_dev", c_void_p), ("ln_u_dev", c_void_p), ("ln_eps", c_float), ("row_sums_dev", c_void_p),
        ("gn_sums_dev", c_void_p), ("gn_cpg", c_int), ("gn_groups", c_int), ("gn_rows_per_img", c_int64),
        ("in_f16", c_int), ("res_f16", c_int),
        ("k_split_ws_dev", c_void_p), ("k_split_ws_floats", c_int64), ("k_split_cnt_dev", c_void_p),
        ("k_split_cnt_len", c_int),
    ]


class ResizeSrc(ctypes.Structure):
    _fields_ = [("ptr_dev", c_void_p), ("h", c_int), ("w", c_int), ("C", c_int), ("c_off", c_int)]


class UNetArch(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int), ("out_channels", c_int), ("num_levels", c_int),
        ("block_out_channels", c_int * GDF_MAX_LEVELS), ("layers_per_block", c_int),
        ("down_has_attn", c_int * GDF_MAX_LEVELS), ("up_has_attn", c_int * GDF_MAX_LEVELS),
        ("transformer_depth", c_int * GDF_MAX_LEVELS), ("num_heads", c_int * GDF_MAX_LEVELS),
        ("cross_attention_dim", c_int), ("use_linear_projection", c_int),
        ("addition_time_embed_dim", c_int), ("projection_class_embeddings_input_dim", c_int),
        ("norm_num_groups", c_int), ("norm_eps", c_float),
    ]


class VaeArch(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int), ("latent_channels", c_int), ("num_levels", c_int),
        ("block_out_channels", c_int * GDF_MAX_LEVELS), ("layers_per_block", c_int),
        ("norm_num_groups", c_int), ("norm_eps", c_float), ("scaling_factor", c_float), ("shift_factor", c_float),
    ]


class DitArch(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int), ("out_channels", c_int), ("patch_size", c_int), ("num_layers", c_int),
        ("num_heads", c_int), ("head_dim", c_int), ("caption_channels", c_int), ("norm_eps", c_float),
    ]


class FluxArch(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int), ("num_layers", c_int), ("num_single_layers", c_int), ("num_heads", c_int),
        ("head_dim", c_int), ("joint_attention_dim", c_int), ("pooled_projection_dim", c_int),
        ("guidance_embeds", c_int),
    ]


class Slot(ctypes.Structure):
    _fields_ = [("offset_bytes", c_int64), ("channels", c_int), ("height", c_int), ("width", c_int),
                ("order", c_int)]


class GdfError(RuntimeError):
    pass


_lib = None


def load():
    """Load libgdf_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdfError(
            "libgdf_b200.so not found at %s - build it with `python -m generic_diffusion_feature_b200.build` "
            "(there is no CPU / PyTorch fallback for the extraction path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.gdf_last_error.restype = ctypes.c_char_p
    lib.gdf_op_groupnorm_workspace_floats.restype = c_int64
    lib.gdf_op_groupnorm_workspace_floats.argtypes = [c_int, c_int]
    P = c_void_p
    lib.gdf_op_linear.argtypes = [P, c_int64, c_int, c_int, P, c_int, c_int, ctypes.POINTER(Epilogue), c_int,
                                  c_int64, c_int64, c_int, P]
    lib.gdf_op_conv3x3.argtypes = [P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, ctypes.POINTER(Epilogue),
                                   c_int, P]
    lib.gdf_op_pack_conv_weight.argtypes = [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]
    lib.gdf_op_conv_in.argtypes = [P, P, P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, P]
    lib.gdf_op_pack_conv_weight_f16.argtypes = [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]
    lib.gdf_op_groupnorm.argtypes = [P, P, P, P, c_int, c_int, c_int, c_int, c_float, c_int, P, P]
    lib.gdf_op_layernorm.argtypes = [P, P, P, P, c_int64, c_int, c_float, P, P, c_int, P]
    lib.gdf_op_attention.argtypes = [P, c_int, P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_float, c_int, P]
    lib.gdf_op_attention_bias.argtypes = [P, c_int, P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_float, P, P]
    lib.gdf_op_softmax_rows.argtypes = [P, c_int64, c_int, c_int, P]
    lib.gdf_debug_attention_trace.argtypes = [P, c_int]
    lib.gdf_op_upsample_nearest2x.argtypes = [P, P, c_int, c_int, c_int, c_int, P]
    lib.gdf_op_im2col_small.argtypes = [P, P, P, c_int, c_int, c_int, c_int, P]
    lib.gdf_op_qsample.argtypes = [P, P, P, c_float, c_float, c_float, c_float, P, P, P, c_int, c_int, P]
    lib.gdf_op_cast_f32_to_bf16.argtypes = [P, P, c_int64, P]
    lib.gdf_op_resize_concat.argtypes = [ctypes.POINTER(ResizeSrc), c_int, c_int, c_int, c_int, c_int, P, P, P, P]
    lib.gdf_op_avgpool_nhwc.argtypes = [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]
    lib.gdf_correspond_workspace_floats.restype = c_int64
    lib.gdf_correspond_workspace_floats.argtypes = [c_int, c_int, c_int]
    lib.gdf_correspond.argtypes = [P, P, c_int, c_int, c_int, P, c_int, P, P, P]
    if hasattr(lib, "gdf_create"):
        lib.gdf_create.argtypes = [ctypes.POINTER(UNetArch), ctypes.POINTER(VaeArch), c_int, ctypes.POINTER(P)]
        lib.gdf_create_dit.argtypes = [ctypes.POINTER(DitArch), ctypes.POINTER(VaeArch), c_int, ctypes.POINTER(P)]
        lib.gdf_denoise_capture_dit.argtypes = [P, c_float, P, c_int, P, P, c_int64, P, P]
        lib.gdf_create_flux.argtypes = [ctypes.POINTER(FluxArch), ctypes.POINTER(VaeArch), c_int, ctypes.POINTER(P)]
        lib.gdf_denoise_capture_flux.argtypes = [P, c_float, c_float, P, c_int, P, P, P, P, c_int64, P, P]
        lib.gdf_destroy.argtypes = [P]
        lib.gdf_load_weights.argtypes = [P, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(P),
                                         ctypes.POINTER(c_int64), ctypes.POINTER(c_int), c_int, P]
        lib.gdf_finalize_weights.argtypes = [P, P]
        lib.gdf_plan.argtypes = [P, ctypes.POINTER(ctypes.c_char_p), c_int, c_int, c_int, ctypes.POINTER(Slot),
                                 ctypes.POINTER(c_int64)]
        lib.gdf_encode_noise.argtypes = [P, P, P, P, c_float, c_float, c_float, P, P]
        lib.gdf_encode_latents.argtypes = [P, P, P, c_float, c_float, c_float, P, P]
        lib.gdf_plan_decoder.argtypes = [P]
        lib.gdf_decode_latents.argtypes = [P, P, c_float, P, c_float, P, P]
        lib.gdf_denoise_capture.argtypes = [P, c_float, P, c_int, P, P, P, c_int64, P, P]
        lib.gdf_control_residual_shapes.argtypes = [P, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int]
        lib.gdf_set_control_residuals.argtypes = [P, ctypes.POINTER(P), c_int, P]
        lib.gdf_plan_generation.argtypes = [P]
        lib.gdf_plan_generation.restype = ctypes.c_uint64
        lib.gdf_set_ctx_len.argtypes = [P, c_int]
        lib.gdf_num_launches.argtypes = [P]
        lib.gdf_workspace_bytes.argtypes = [P]
        lib.gdf_workspace_bytes.restype = c_int64
        lib.gdf_profile.argtypes = [P, c_int]
        lib.gdf_profile_dump.argtypes = [P, ctypes.c_char_p]
        lib.gdf_profile_read.argtypes = [P, ctypes.POINTER(c_float), ctypes.POINTER(ctypes.c_double),
                                         ctypes.POINTER(c_int)]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GdfError("gdf error %d: %s" % (rc, load().gdf_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda, "gdf kernels take device tensors only"
    return c_void_p(t.data_ptr())


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)
