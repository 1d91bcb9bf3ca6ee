"""Import shim that EXECUTES the reference's vendored diffusers modules in a container without diffusers.

The reference patches files of diffusers==0.32.2 (feature/diffusers/models/{resnet,attention,attention_processor,
downsampling,upsampling}.py, transformers/transformer_2d.py, unet/unet_2d_condition.py). They use package-relative
imports, so this module builds a fake package `refdiffusers` in sys.modules whose un-vendored pieces are tiny
stand-ins written here (utils, configuration_utils, loaders, activations, embeddings, normalization,
modeling_utils, modeling_outputs, unet_2d_blocks) and whose vendored pieces are loaded FROM THEIR SOURCE FILES
under /root/reference at run time (nothing is copied into this repository).

What this pins: every line of the vendored forward passes (ResnetBlock2D, Downsample2D, Upsample2D, Attention +
AttnProcessor2_0, FeedForward, BasicTransformerBlock, Transformer2DModel, UNet2DConditionModel.forward incl. the
time / text_time embedding plumbing and all 27 gather call sites) and the real FeatureStore /
prepare_feature_extractor / correspondence_utils. What it does NOT pin (restated from the published diffusers
0.32.2 semantics): the block wiring in unet_2d_blocks, Timesteps / TimestepEmbedding,
GEGLU / get_activation. Only tools/make_golden.py uses this; it cannot run on the GPU box (no /root/reference).
"""
import dataclasses
import functools
import importlib.util
import inspect
import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("GDF_REFERENCE", "/root/reference")
PKG = "refdiffusers"


def _mod(name, is_pkg=False):
    m = types.ModuleType(name)
    if is_pkg:
        m.__path__ = []
    m.__package__ = name if is_pkg else name.rpartition(".")[0]
    sys.modules[name] = m
    return m


def _load_ref(name, relpath):
    path = os.path.join(REF, "feature", "diffusers", relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    m.__package__ = name.rpartition(".")[0]
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        cfg = {k: v.default for k, v in list(sig.parameters.items())[1:] if v.default is not inspect._empty}
        names = list(sig.parameters.keys())[1:]
        for n, a in zip(names, args):
            cfg[n] = a
        cfg.update(kwargs)
        self._internal_dict = _Cfg(cfg)
        init(self, *args, **kwargs)
    return inner


def install():
    if PKG + ".models.unet.unet_2d_condition" in sys.modules:
        return sys.modules[PKG]
    root = _mod(PKG, True)

    # ---- utils
    utils = _mod(PKG + ".utils", True)

    def deprecate(*a, **k):
        return None

    class _Logger:
        def warning(self, *a, **k):
            pass
        info = debug = warning_once = warning

    class _logging:
        @staticmethod
        def get_logger(name=None):
            return _Logger()

    def is_torch_version(op, ver):
        from packaging import version
        cur = version.parse(torch.__version__.split("+")[0])
        v = version.parse(ver)
        return {"<": cur < v, "<=": cur <= v, ">": cur > v, ">=": cur >= v, "==": cur == v}[op]

    @dataclasses.dataclass
    class BaseOutput:
        pass

    utils.deprecate = deprecate
    utils.logging = _logging
    utils.is_torch_version = is_torch_version
    utils.is_torch_xla_available = lambda: False
    utils.USE_PEFT_BACKEND = False
    utils.BaseOutput = BaseOutput
    utils.scale_lora_layers = lambda *a, **k: None
    utils.unscale_lora_layers = lambda *a, **k: None
    tu = _mod(PKG + ".utils.torch_utils")
    tu.maybe_allow_in_graph = lambda cls: cls
    tu.is_torch_version = is_torch_version
    iu = _mod(PKG + ".utils.import_utils")
    iu.is_torch_npu_available = lambda: False
    iu.is_torch_xla_version = lambda *a: False
    iu.is_xformers_available = lambda: False
    iu.is_torch_version = is_torch_version
    utils.torch_utils, utils.import_utils = tu, iu

    ip = _mod(PKG + ".image_processor")
    ip.IPAdapterMaskProcessor = type("IPAdapterMaskProcessor", (), {})

    cu = _mod(PKG + ".configuration_utils")

    class ConfigMixin:
        @property
        def config(self):
            return self._internal_dict

        def register_to_config(self, **kw):
            self._internal_dict.update(kw)

    cu.ConfigMixin = ConfigMixin
    cu.LegacyConfigMixin = ConfigMixin
    cu.register_to_config = _register_to_config

    ld = _mod(PKG + ".loaders", True)
    ld.PeftAdapterMixin = type("PeftAdapterMixin", (), {})
    ld.UNet2DConditionLoadersMixin = type("UNet2DConditionLoadersMixin", (), {})
    sf = _mod(PKG + ".loaders.single_file_model")
    sf.FromOriginalModelMixin = type("FromOriginalModelMixin", (), {})

    models = _mod(PKG + ".models", True)

    # ---- un-vendored: activations [diffusers 0.32.2, restated]
    act = _mod(PKG + ".models.activations")

    def get_activation(name):
        return {"swish": nn.SiLU(), "silu": nn.SiLU(), "mish": nn.Mish(), "gelu": nn.GELU(), "relu": nn.ReLU()}[name]

    class GEGLU(nn.Module):
        def __init__(self, dim_in, dim_out, bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out * 2, bias=bias)

        def forward(self, hidden_states, *a, **k):
            hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
            return hidden_states * F.gelu(gate)

    class GELU(nn.Module):
        def __init__(self, dim_in, dim_out, approximate="none", bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out, bias=bias)
            self.approximate = approximate

        def forward(self, x):
            return F.gelu(self.proj(x), approximate=self.approximate)

    act.get_activation, act.GEGLU, act.GELU = get_activation, GEGLU, GELU
    for n in ("ApproximateGELU", "FP32SiLU", "LinearActivation", "SwiGLU"):
        setattr(act, n, type(n, (nn.Module,), {}))

    # ---- un-vendored: normalization (only names are needed on the UNet path)
    nm = _mod(PKG + ".models.normalization")
    for n in ("AdaGroupNorm", "AdaLayerNorm", "AdaLayerNormContinuous", "AdaLayerNormZero", "RMSNorm",
              "SD35AdaLayerNormZeroX", "AdaLayerNormSingle", "FP32LayerNorm", "LpNorm"):
        setattr(nm, n, type(n, (nn.Module,), {}))

    # ---- un-vendored: embeddings [diffusers 0.32.2, restated]
    emb = _mod(PKG + ".models.embeddings")

    class Timesteps(nn.Module):
        def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale=1):
            super().__init__()
            self.num_channels, self.flip, self.shift, self.scale = num_channels, flip_sin_to_cos, \
                downscale_freq_shift, scale

        def forward(self, timesteps):
            half = self.num_channels // 2
            exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
            exponent = exponent / (half - self.shift)
            e = timesteps[:, None].float() * torch.exp(exponent)[None, :] * self.scale
            e = torch.cat([torch.sin(e), torch.cos(e)], dim=-1)
            if self.flip:
                e = torch.cat([e[:, half:], e[:, :half]], dim=-1)
            return e

    class TimestepEmbedding(nn.Module):
        def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None,
                     cond_proj_dim=None, sample_proj_bias=True):
            super().__init__()
            self.linear_1 = nn.Linear(in_channels, time_embed_dim, sample_proj_bias)
            self.act = get_activation(act_fn)
            self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim, sample_proj_bias)

        def forward(self, sample, condition=None):
            return self.linear_2(self.act(self.linear_1(sample)))

    emb.Timesteps, emb.TimestepEmbedding = Timesteps, TimestepEmbedding
    for n in ("SinusoidalPositionalEmbedding", "ImagePositionalEmbeddings", "PatchEmbed", "PixArtAlphaTextProjection",
              "GaussianFourierProjection", "GLIGENTextBoundingboxProjection", "ImageHintTimeEmbedding",
              "ImageProjection", "ImageTimeEmbedding", "TextImageProjection", "TextImageTimeEmbedding",
              "TextTimeEmbedding"):
        setattr(emb, n, type(n, (nn.Module,), {}))

    mo = _mod(PKG + ".models.modeling_outputs")

    @dataclasses.dataclass
    class Transformer2DModelOutput:
        sample: torch.Tensor = None

    mo.Transformer2DModelOutput = Transformer2DModelOutput
    mu = _mod(PKG + ".models.modeling_utils")

    class ModelMixin(nn.Module):
        pass

    mu.ModelMixin = ModelMixin
    mu.LegacyModelMixin = ModelMixin

    # ---- vendored reference files, loaded from source in dependency order
    up = _load_ref(PKG + ".models.upsampling", "models/upsampling.py")
    dn = _load_ref(PKG + ".models.downsampling", "models/downsampling.py")
    ap = _load_ref(PKG + ".models.attention_processor", "models/attention_processor.py")
    rn = _load_ref(PKG + ".models.resnet", "models/resnet.py")
    at = _load_ref(PKG + ".models.attention", "models/attention.py")
    _mod(PKG + ".models.transformers", True)
    t2 = _load_ref(PKG + ".models.transformers.transformer_2d", "models/transformers/transformer_2d.py")
    _mod(PKG + ".models.unet", True)

    # ---- un-vendored: unet_2d_blocks wiring [diffusers 0.32.2, restated] on top of the vendored modules
    blk = _mod(PKG + ".models.unet.unet_2d_blocks")

    def _resnet(cin, cout, temb, eps, groups):
        return rn.ResnetBlock2D(in_channels=cin, out_channels=cout, temb_channels=temb, eps=eps, groups=groups,
                                dropout=0.0, time_embedding_norm="default", non_linearity="swish",
                                output_scale_factor=1.0, pre_norm=True)

    def _vit(heads, ch, depth, ctx, groups, linear):
        return t2.Transformer2DModel(heads, ch // heads, in_channels=ch, num_layers=depth, cross_attention_dim=ctx,
                                     norm_num_groups=groups, use_linear_projection=linear,
                                     only_cross_attention=False, upcast_attention=False, attention_type="default")

    class DownBlock(nn.Module):
        def __init__(self, has_attn, in_channels, out_channels, temb_channels, num_layers, depth, eps, groups, heads,
                     ctx, linear, add_downsample, pad):
            super().__init__()
            self.has_cross_attention = has_attn
            self.resnets = nn.ModuleList([_resnet(in_channels if i == 0 else out_channels, out_channels,
                                                  temb_channels, eps, groups) for i in range(num_layers)])
            if has_attn:
                self.attentions = nn.ModuleList([_vit(heads, out_channels, depth, ctx, groups, linear)
                                                 for _ in range(num_layers)])
            self.downsamplers = nn.ModuleList([dn.Downsample2D(out_channels, use_conv=True,
                                                               out_channels=out_channels, padding=pad, name="op")]) \
                if add_downsample else None

        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                    cross_attention_kwargs=None, encoder_attention_mask=None, **kw):
            out = ()
            for i, r in enumerate(self.resnets):
                hidden_states = r(hidden_states, temb)
                if self.has_cross_attention:
                    hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                       cross_attention_kwargs=cross_attention_kwargs,
                                                       attention_mask=attention_mask,
                                                       encoder_attention_mask=encoder_attention_mask,
                                                       return_dict=False)[0]
                out = out + (hidden_states,)
            if self.downsamplers is not None:
                for d in self.downsamplers:
                    hidden_states = d(hidden_states)
                out = out + (hidden_states,)
            return hidden_states, out

    class MidBlock(nn.Module):
        def __init__(self, ch, temb, eps, groups, depth, heads, ctx, linear):
            super().__init__()
            self.has_cross_attention = True
            self.resnets = nn.ModuleList([_resnet(ch, ch, temb, eps, groups), _resnet(ch, ch, temb, eps, groups)])
            self.attentions = nn.ModuleList([_vit(heads, ch, depth, ctx, groups, linear)])

        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                    cross_attention_kwargs=None, encoder_attention_mask=None):
            hidden_states = self.resnets[0](hidden_states, temb)
            for a, r in zip(self.attentions, self.resnets[1:]):
                hidden_states = a(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                  cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                  encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
                hidden_states = r(hidden_states, temb)
            return hidden_states

    class UpBlock(nn.Module):
        def __init__(self, has_attn, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, depth,
                     eps, groups, heads, ctx, linear, add_upsample):
            super().__init__()
            self.has_cross_attention = has_attn
            rs = []
            for i in range(num_layers):
                res_skip = in_channels if (i == num_layers - 1) else out_channels
                rin = prev_output_channel if i == 0 else out_channels
                rs.append(_resnet(rin + res_skip, out_channels, temb_channels, eps, groups))
            self.resnets = nn.ModuleList(rs)
            if has_attn:
                self.attentions = nn.ModuleList([_vit(heads, out_channels, depth, ctx, groups, linear)
                                                 for _ in range(num_layers)])
            self.upsamplers = nn.ModuleList([up.Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
                if add_upsample else None

        def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                    cross_attention_kwargs=None, upsample_size=None, attention_mask=None,
                    encoder_attention_mask=None):
            for i, r in enumerate(self.resnets):
                res = res_hidden_states_tuple[-1]
                res_hidden_states_tuple = res_hidden_states_tuple[:-1]
                hidden_states = torch.cat([hidden_states, res], dim=1)
                hidden_states = r(hidden_states, temb)
                if self.has_cross_attention:
                    hidden_states = self.attentions[i](hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                       cross_attention_kwargs=cross_attention_kwargs,
                                                       attention_mask=attention_mask,
                                                       encoder_attention_mask=encoder_attention_mask,
                                                       return_dict=False)[0]
            if self.upsamplers is not None:
                for u in self.upsamplers:
                    hidden_states = u(hidden_states, upsample_size)
            return hidden_states

    def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                       resnet_eps, resnet_act_fn, transformer_layers_per_block=1, num_attention_heads=None,
                       resnet_groups=None, cross_attention_dim=None, downsample_padding=None,
                       use_linear_projection=False, **kw):
        return DownBlock(down_block_type == "CrossAttnDownBlock2D", in_channels, out_channels, temb_channels,
                         num_layers, transformer_layers_per_block, resnet_eps, resnet_groups, num_attention_heads,
                         cross_attention_dim, use_linear_projection, add_downsample, downsample_padding)

    def get_mid_block(mid_block_type, temb_channels, in_channels, resnet_eps, resnet_act_fn, resnet_groups,
                      transformer_layers_per_block=1, num_attention_heads=None, cross_attention_dim=None,
                      use_linear_projection=False, **kw):
        return MidBlock(in_channels, temb_channels, resnet_eps, resnet_groups, transformer_layers_per_block,
                        num_attention_heads, cross_attention_dim, use_linear_projection)

    def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, temb_channels,
                     add_upsample, resnet_eps, resnet_act_fn, transformer_layers_per_block=1, num_attention_heads=None,
                     resnet_groups=None, cross_attention_dim=None, use_linear_projection=False, **kw):
        return UpBlock(up_block_type == "CrossAttnUpBlock2D", in_channels, out_channels, prev_output_channel,
                       temb_channels, num_layers, transformer_layers_per_block, resnet_eps, resnet_groups,
                       num_attention_heads, cross_attention_dim, use_linear_projection, add_upsample)

    blk.get_down_block, blk.get_mid_block, blk.get_up_block = get_down_block, get_mid_block, get_up_block
    un = _load_ref(PKG + ".models.unet.unet_2d_condition", "models/unet/unet_2d_condition.py")
    root.resnet, root.attention, root.attention_processor = rn, at, ap
    root.transformer_2d, root.unet_2d_condition, root.downsampling, root.upsampling = t2, un, dn, up
    return root


def install_flux():
    """Adds what the reference's vendored transformers/transformer_flux.py imports and loads it FROM ITS SOURCE FILE.
    Vendored (pinned): FluxTransformer2DModel (ctor + forward), FluxTransformerBlock, FluxSingleTransformerBlock,
    Attention ctor + FluxAttnProcessor2_0, FeedForward. Un-vendored, restated here from the published diffusers 0.32.2
    semantics (NOT pinned): normalization.{RMSNorm, AdaLayerNormZero, AdaLayerNormZeroSingle, AdaLayerNormContinuous},
    embeddings.{FluxPosEmbed, apply_rotary_emb, CombinedTimestep(Guidance)TextProjEmbeddings,
    PixArtAlphaTextProjection}."""
    root = install()
    if PKG + ".models.transformers.transformer_flux" in sys.modules:
        return root
    nm = sys.modules[PKG + ".models.normalization"]
    emb = sys.modules[PKG + ".models.embeddings"]
    ld = sys.modules[PKG + ".loaders"]
    ld.FluxTransformer2DLoadersMixin = type("FluxTransformer2DLoadersMixin", (), {})
    ld.FromOriginalModelMixin = sys.modules[PKG + ".loaders.single_file_model"].FromOriginalModelMixin

    class RMSNorm(nn.Module):
        def __init__(self, dim, eps, elementwise_affine=True, bias=False):
            super().__init__()
            self.eps = eps
            self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None

        def forward(self, hidden_states):
            input_dtype = hidden_states.dtype
            variance = hidden_states.to(torch.float32).pow(2).mean(-1, keepdim=True)
            hidden_states = hidden_states * torch.rsqrt(variance + self.eps)
            if self.weight is not None:
                hidden_states = hidden_states * self.weight
            return hidden_states.to(input_dtype)

    class AdaLayerNormZero(nn.Module):
        def __init__(self, embedding_dim, num_embeddings=None, norm_type="layer_norm", bias=True):
            super().__init__()
            self.emb = None
            self.silu = nn.SiLU()
            self.linear = nn.Linear(embedding_dim, 6 * embedding_dim, bias=bias)
            self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

        def forward(self, x, timestep=None, class_labels=None, hidden_dtype=None, emb=None):
            emb = self.linear(self.silu(emb))
            shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
            x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
            return x, gate_msa, shift_mlp, scale_mlp, gate_mlp

    class AdaLayerNormZeroSingle(nn.Module):
        def __init__(self, embedding_dim, norm_type="layer_norm", bias=True):
            super().__init__()
            self.silu = nn.SiLU()
            self.linear = nn.Linear(embedding_dim, 3 * embedding_dim, bias=bias)
            self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

        def forward(self, x, emb=None):
            emb = self.linear(self.silu(emb))
            shift_msa, scale_msa, gate_msa = emb.chunk(3, dim=1)
            x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
            return x, gate_msa

    class AdaLayerNormContinuous(nn.Module):
        def __init__(self, embedding_dim, conditioning_embedding_dim, elementwise_affine=True, eps=1e-5, bias=True,
                     norm_type="layer_norm"):
            super().__init__()
            self.silu = nn.SiLU()
            self.linear = nn.Linear(conditioning_embedding_dim, embedding_dim * 2, bias=bias)
            self.norm = nn.LayerNorm(embedding_dim, eps, elementwise_affine, bias)

        def forward(self, x, conditioning_embedding):
            emb = self.linear(self.silu(conditioning_embedding).to(x.dtype))
            scale, shift = torch.chunk(emb, 2, dim=1)
            return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]

    nm.RMSNorm, nm.AdaLayerNormZero, nm.AdaLayerNormZeroSingle = RMSNorm, AdaLayerNormZero, AdaLayerNormZeroSingle
    nm.AdaLayerNormContinuous = AdaLayerNormContinuous

    def get_1d_rotary_pos_embed(dim, pos, theta=10000.0):
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: (dim // 2)] / dim))
        freqs = torch.outer(pos, freqs)
        return freqs.cos().repeat_interleave(2, dim=1).float(), freqs.sin().repeat_interleave(2, dim=1).float()

    class FluxPosEmbed(nn.Module):
        def __init__(self, theta, axes_dim):
            super().__init__()
            self.theta, self.axes_dim = theta, axes_dim

        def forward(self, ids):
            cos_out, sin_out = [], []
            pos = ids.float().double()
            for i in range(ids.shape[-1]):
                c, s_ = get_1d_rotary_pos_embed(self.axes_dim[i], pos[:, i], self.theta)
                cos_out.append(c)
                sin_out.append(s_)
            return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)

    def apply_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1):
        cos, sin = freqs_cis
        cos, sin = cos[None, None], sin[None, None]
        x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
        x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
        return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)

    class PixArtAlphaTextProjection(nn.Module):
        def __init__(self, in_features, hidden_size, out_features=None, act_fn="gelu_tanh"):
            super().__init__()
            self.linear_1 = nn.Linear(in_features, hidden_size)
            self.act_1 = {"gelu_tanh": nn.GELU(approximate="tanh"), "silu": nn.SiLU()}[act_fn]
            self.linear_2 = nn.Linear(hidden_size, out_features or hidden_size)

        def forward(self, caption):
            return self.linear_2(self.act_1(self.linear_1(caption)))

    class CombinedTimestepTextProjEmbeddings(nn.Module):
        def __init__(self, embedding_dim, pooled_projection_dim):
            super().__init__()
            self.time_proj = emb.Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)
            self.timestep_embedder = emb.TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
            self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim, act_fn="silu")

        def forward(self, timestep, pooled_projection):
            t = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled_projection.dtype))
            return t + self.text_embedder(pooled_projection)

    class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
        def __init__(self, embedding_dim, pooled_projection_dim):
            super().__init__()
            self.time_proj = emb.Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)
            self.timestep_embedder = emb.TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
            self.guidance_embedder = emb.TimestepEmbedding(in_channels=256, time_embed_dim=embedding_dim)
            self.text_embedder = PixArtAlphaTextProjection(pooled_projection_dim, embedding_dim, act_fn="silu")

        def forward(self, timestep, guidance, pooled_projection):
            t = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled_projection.dtype))
            g = self.guidance_embedder(self.time_proj(guidance).to(dtype=pooled_projection.dtype))
            return t + g + self.text_embedder(pooled_projection)

    emb.FluxPosEmbed, emb.apply_rotary_emb = FluxPosEmbed, apply_rotary_emb
    emb.PixArtAlphaTextProjection = PixArtAlphaTextProjection
    emb.CombinedTimestepTextProjEmbeddings = CombinedTimestepTextProjEmbeddings
    emb.CombinedTimestepGuidanceTextProjEmbeddings = CombinedTimestepGuidanceTextProjEmbeddings
    root.transformer_flux = _load_ref(PKG + ".models.transformers.transformer_flux",
                                      "models/transformers/transformer_flux.py")
    return root


def build_reference_flux(cfg):
    """The reference's (vendored) FluxTransformer2DModel for one of the oracle-style config dicts."""
    root = install_flux()
    return root.transformer_flux.FluxTransformer2DModel(
        patch_size=1, in_channels=cfg["in_ch"], num_layers=cfg["layers"], num_single_layers=cfg["single_layers"],
        attention_head_dim=cfg["head_dim"], num_attention_heads=cfg["heads"], joint_attention_dim=cfg["joint_dim"],
        pooled_projection_dim=cfg["pooled_dim"], guidance_embeds=cfg["guidance_embeds"],
        axes_dims_rope=tuple(cfg["axes_dims_rope"]))


def build_reference_unet(cfg):
    """The reference's (vendored) UNet2DConditionModel for one of the oracle-style config dicts."""
    root = install()
    n = len(cfg["block_out"])
    down = tuple("CrossAttnDownBlock2D" if a else "DownBlock2D" for a in cfg["down_attn"])
    upb = tuple("CrossAttnUpBlock2D" if a else "UpBlock2D" for a in cfg["up_attn"])
    kw = dict(sample_size=None, in_channels=4, out_channels=4, down_block_types=down, up_block_types=upb,
              block_out_channels=tuple(cfg["block_out"]), layers_per_block=2, cross_attention_dim=cfg["ctx_dim"],
              transformer_layers_per_block=tuple(cfg["depth"]), attention_head_dim=tuple(cfg["heads"]),
              use_linear_projection=cfg["linear_proj"], norm_eps=cfg["eps"], norm_num_groups=32)
    if cfg["add_time_dim"]:
        kw.update(addition_embed_type="text_time", addition_time_embed_dim=cfg["add_time_dim"],
                  projection_class_embeddings_input_dim=cfg["add_in"])
    assert n == len(down)
    return root.unet_2d_condition.UNet2DConditionModel(**kw)


def load_reference_feature_extractor():
    """The reference's own FeatureStore / prepare_feature_extractor (feature/components/feature_extractor.py)."""
    path = os.path.join(REF, "feature", "components", "feature_extractor.py")
    spec = importlib.util.spec_from_file_location("ref_feature_extractor", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference_attention_store():
    """feature/components/attention.py (AttentionStore, AttnStoreProcessor, register_attention_store) executed from its
    source file; its `from diffusers.models.attention_processor import ...` resolves to the vendored module loaded by
    install()."""
    root = install()
    for alias, target in (("diffusers", PKG), ("diffusers.models", PKG + ".models"),
                          ("diffusers.models.attention_processor", PKG + ".models.attention_processor"),
                          ("diffusers.models.embeddings", PKG + ".models.embeddings")):   # FluxAttnStoreProcessor: rotary
        if target in sys.modules:
            sys.modules.setdefault(alias, sys.modules[target])
    if not hasattr(root.attention_processor, "AttnProcessor"):
        raise RuntimeError("vendored attention_processor has no AttnProcessor")
    path = os.path.join(REF, "feature", "components", "attention.py")
    spec = importlib.util.spec_from_file_location("ref_components_attention", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference_correspondence_utils():
    """correspondence/correspondence/correspondence_utils.py with a stub for its (unused here) matplotlib import."""
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    sys.modules["matplotlib.patches"].ConnectionPatch = object
    sys.modules["matplotlib.colors"].ListedColormap = object
    path = os.path.join(REF, "correspondence", "correspondence", "correspondence_utils.py")
    spec = importlib.util.spec_from_file_location("ref_correspondence_utils", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_reference_segmentor():
    """segmentation/models/diffusion_segmentor.py (ResBlock, MultiRes, DiffusionSegmentor.extract_feat) with stubs for
    its mmseg / mmengine imports and its relative `.base` import: only the registry decorator, the type aliases and a
    BaseSegmentor that is a plain nn.Module are needed to execute ResBlock.forward and extract_feat."""
    import torch.nn as nn
    saved = {n: sys.modules.get(n) for n in ("diffusion_feature",)}
    for n in ("mmengine", "mmengine.logging", "mmseg", "mmseg.registry", "mmseg.utils", "ref_segmodels",
              "ref_segmodels.base"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    sys.modules["mmengine.logging"].print_log = print

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

        def build(self, cfg):
            raise RuntimeError("stub registry: nothing to build")

    sys.modules["mmseg.registry"].MODELS = _Registry()
    for n in ("ConfigType", "OptConfigType", "OptMultiConfig", "OptSampleList", "SampleList"):
        setattr(sys.modules["mmseg.utils"], n, object)
    sys.modules["mmseg.utils"].add_prefix = lambda d, p: d
    sys.modules["ref_segmodels.base"].BaseSegmentor = nn.Module
    stub = types.ModuleType("diffusion_feature")
    stub.FeatureExtractor = object
    sys.modules["diffusion_feature"] = stub
    path = os.path.join(REF, "segmentation", "models", "diffusion_segmentor.py")
    spec = importlib.util.spec_from_file_location("ref_segmodels.diffusion_segmentor", path)
    m = importlib.util.module_from_spec(spec)
    m.__package__ = "ref_segmodels"
    try:
        spec.loader.exec_module(m)

        class _TorchOnCpu:
            """The several-extractors branch of extract_feat moves every map `.to(torch.device("cuda:0"))`
            (diffusion_segmentor.py:267): in this GPU-less container that one name resolves to the CPU."""
            def __getattr__(self, n):
                return getattr(torch, n)

            def device(self, *a, **k):
                return torch.device("cpu")

        m.torch = _TorchOnCpu()
    finally:
        if saved["diffusion_feature"] is None:
            del sys.modules["diffusion_feature"]
        else:
            sys.modules["diffusion_feature"] = saved["diffusion_feature"]
    return m
