"""Model factory of the B200 path: host-side mirror of feature/components/models.py (reference).

`get_diffusion_model(version, dtype, ...)` returns a `B200Pipe`, the object `FeatureExtractor` drives instead
of a diffusers pipeline: it owns a libgdf_b200 handle (packed bf16 weights on the device) and the architecture /
scheduler constants the reference reads from `pipe.unet.config`, `pipe.vae.config` and `pipe.scheduler`.

Version strings and error behaviour follow models.py:10-16,173-174: dtype must be 'float16' or 'float32'
(both select the same bf16-compute kernels; the returned features are fp16 either way, as FeatureStore.store
casts them, feature_extractor.py:59-60), unknown versions raise NotImplementedError.
"""
import ctypes
import zlib

import torch

from .. import _lib
from .._lib import DitArch, FluxArch, UNetArch, VaeArch, check

# SURVEY.md Appendix B (diffusers configs of the checkpoints models.py:18-70 names)
UNET_CONFIGS = {
    "xl": dict(block_out=(320, 640, 1280), down_attn=(0, 1, 1), up_attn=(1, 1, 0), depth=(1, 2, 10),
               heads=(5, 10, 20), ctx_dim=2048, linear_proj=True, add_time_dim=256, add_in=2816, eps=1e-5),
    "1-5": dict(block_out=(320, 640, 1280, 1280), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
                heads=(8, 8, 8, 8), ctx_dim=768, linear_proj=False, add_time_dim=0, add_in=0, eps=1e-5),
    "2-1": dict(block_out=(320, 640, 1280, 1280), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
                heads=(5, 10, 20, 20), ctx_dim=1024, linear_proj=True, add_time_dim=0, add_in=0, eps=1e-5),
}
UNET_CONFIGS["pgv2"] = UNET_CONFIGS["xl"]
VAE_CONFIGS = {
    "xl": dict(block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6, scaling_factor=0.13025),
    "pgv2": dict(block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6, scaling_factor=0.13025),
    "1-5": dict(block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6, scaling_factor=0.18215),
    "2-1": dict(block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6, scaling_factor=0.18215),
}
# [PixArt-alpha/PixArt-Sigma-XL-2-{1024,512}-MS transformer/config.json, from memory; SURVEY.md row a16]
DIT_CONFIGS = {
    "pixart-sigma": dict(layers=28, heads=16, head_dim=72, in_ch=4, out_ch=8, patch=2, caption_dim=4096,
                         sample_size=128, interpolation_scale=2.0, eps=1e-6),
    "pixart-sigma-512": dict(layers=28, heads=16, head_dim=72, in_ch=4, out_ch=8, patch=2, caption_dim=4096,
                             sample_size=64, interpolation_scale=1.0, eps=1e-6),
}
# [PixArt-alpha/PixArt-XL-2-512x512 transformer/config.json, from memory]: the same 28 x (16 x 72) AdaLN-single blocks as
# Sigma-512 (sample_size 64, interpolation_scale 1, caption_channels 4096; no resolution / aspect-ratio conditioning at
# 512: `use_additional_conditions` is sample_size == 128 only); prompts are 120 T5 tokens (diffusion_feature.py:195-205
# passes whatever encode_prompt returned), the VAE is the SD one (sd-vae-ft-ema, scaling_factor 0.18215).
DIT_CONFIGS["pixart-alpha"] = dict(DIT_CONFIGS["pixart-sigma-512"], prompt_len=120)
VAE_CONFIGS["pixart-sigma"] = VAE_CONFIGS["xl"]          # PixArt-Sigma ships the SDXL VAE
VAE_CONFIGS["pixart-sigma-512"] = VAE_CONFIGS["xl"]
VAE_CONFIGS["pixart-alpha"] = VAE_CONFIGS["1-5"]
# [black-forest-labs/FLUX.1-dev transformer/config.json = the FluxTransformer2DModel defaults the reference vendors at
# transformers/transformer_flux.py:253-266 + guidance_embeds true; vae/config.json, from memory: 16 latent channels,
# no quant_conv, scaling_factor 0.3611, shift_factor 0.1159]
FLUX_CONFIGS = {
    "flux": dict(layers=19, single_layers=38, heads=24, head_dim=128, in_ch=64, joint_dim=4096, pooled_dim=768,
                 guidance_embeds=True, axes_dims_rope=(16, 56, 56), ctx_len=512),
}
VAE_CONFIGS["flux"] = dict(block_out=(128, 256, 512, 512), layers=2, latent=16, eps=1e-6, scaling_factor=0.3611,
                           shift_factor=0.1159, quant_conv=False)
_NOT_BUILT = ("if", "hunyuan")


def flux_param_specs(cfg):
    """(name, shape) of every FluxTransformer2DModel parameter (transformer_flux.py:253-325; Attention ctor
    attention_processor.py:105-297 with qk_norm='rms_norm', added_kv_proj_dim / pre_only), diffusers naming."""
    C = cfg["heads"] * cfg["head_dim"]
    hd = cfg["head_dim"]
    out = []

    def lin(n, o, i):
        out.append((n + ".weight", (o, i)))
        out.append((n + ".bias", (o,)))

    lin("x_embedder", C, cfg["in_ch"])
    lin("context_embedder", C, cfg["joint_dim"])
    embs = ["timestep_embedder"] + (["guidance_embedder"] if cfg["guidance_embeds"] else [])
    for e in embs:
        lin("time_text_embed.%s.linear_1" % e, C, 256)
        lin("time_text_embed.%s.linear_2" % e, C, C)
    lin("time_text_embed.text_embedder.linear_1", C, cfg["pooled_dim"])
    lin("time_text_embed.text_embedder.linear_2", C, C)
    for k in range(cfg["layers"]):
        b = "transformer_blocks.%d" % k
        lin(b + ".norm1.linear", 6 * C, C)
        lin(b + ".norm1_context.linear", 6 * C, C)
        for pn in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            lin("%s.attn.%s" % (b, pn), C, C)
        for nn_ in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            out.append(("%s.attn.%s.weight" % (b, nn_), (hd,)))
        for f in ("ff", "ff_context"):
            lin("%s.%s.net.0.proj" % (b, f), 4 * C, C)
            lin("%s.%s.net.2" % (b, f), C, 4 * C)
    for k in range(cfg["single_layers"]):
        b = "single_transformer_blocks.%d" % k
        lin(b + ".norm.linear", 3 * C, C)
        lin(b + ".proj_mlp", 4 * C, C)
        lin(b + ".proj_out", C, 5 * C)
        for pn in ("to_q", "to_k", "to_v"):
            lin("%s.attn.%s" % (b, pn), C, C)
        for nn_ in ("norm_q", "norm_k"):
            out.append(("%s.attn.%s.weight" % (b, nn_), (hd,)))
    lin("norm_out.linear", 2 * C, C)
    lin("proj_out", cfg["in_ch"], C)
    return out


def flux_rope_tables(cfg, ctx_len, grid):
    """cos / sin tables [ctx_len + grid*grid, head_dim] of FluxPosEmbed for ids = cat(txt_ids (zeros),
    img_ids (0, y, x)) (pipeline_flux_img2img.py:483-494, transformer_flux.py:481-482). [diffusers embeddings.
    FluxPosEmbed / get_1d_rotary_pos_embed(use_real=True, repeat_interleave_real=True, freqs_dtype=float64),
    un-vendored]: per axis, freqs = pos * theta^(-2i/d); cos / sin repeat-interleaved by 2; axes concatenated."""
    import numpy as np
    S = ctx_len + grid * grid
    ids = np.zeros((S, 3), dtype=np.float64)
    yy, xx = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    ids[ctx_len:, 1] = yy.reshape(-1)
    ids[ctx_len:, 2] = xx.reshape(-1)
    cos, sin = [], []
    for a, d in enumerate(cfg["axes_dims_rope"]):
        freqs = 1.0 / (10000.0 ** (np.arange(0, d, 2, dtype=np.float64)[: d // 2] / d))
        ang = np.outer(ids[:, a], freqs)
        cos.append(np.repeat(np.cos(ang), 2, axis=1))
        sin.append(np.repeat(np.sin(ang), 2, axis=1))
    cos = torch.from_numpy(np.concatenate(cos, axis=1)).float()
    sin = torch.from_numpy(np.concatenate(sin, axis=1)).float()
    assert cos.shape[1] == cfg["head_dim"], "axes_dims_rope must sum to head_dim"
    return cos, sin


def dit_param_specs(cfg):
    """(name, shape) of every PixArtTransformer2DModel parameter / persistent buffer, diffusers naming."""
    C = cfg["heads"] * cfg["head_dim"]
    grid = cfg["sample_size"] // cfg["patch"]
    out = []

    def lin(n, o, i):
        out.append((n + ".weight", (o, i)))
        out.append((n + ".bias", (o,)))

    out.append(("pos_embed.proj.weight", (C, cfg["in_ch"], cfg["patch"], cfg["patch"])))
    out.append(("pos_embed.proj.bias", (C,)))
    out.append(("pos_embed.pos_embed", (1, grid * grid, C)))
    lin("adaln_single.emb.timestep_embedder.linear_1", C, 256)
    lin("adaln_single.emb.timestep_embedder.linear_2", C, C)
    lin("adaln_single.linear", 6 * C, C)
    lin("caption_projection.linear_1", C, cfg["caption_dim"])
    lin("caption_projection.linear_2", C, C)
    for k in range(cfg["layers"]):
        b = "transformer_blocks.%d" % k
        out.append((b + ".scale_shift_table", (6, C)))
        for a in ("attn1", "attn2"):
            for pn in ("to_q", "to_k", "to_v", "to_out.0"):
                lin("%s.%s.%s" % (b, a, pn), C, C)
        lin(b + ".ff.net.0.proj", 4 * C, C)
        lin(b + ".ff.net.2", C, 4 * C)
    out.append(("scale_shift_table", (2, C)))
    lin("proj_out", cfg["patch"] * cfg["patch"] * cfg["out_ch"], C)
    return out


def sincos_pos_embed_2d(dim, grid, base_size, interpolation_scale):
    """[diffusers embeddings.get_2d_sincos_pos_embed]: the constant table PatchEmbed registers as the persistent
    buffer `pos_embed` (so a real checkpoint's state_dict carries it); generated here for synthetic weights."""
    import numpy as np
    pos = np.arange(grid, dtype=np.float32).astype(np.float64) / (grid / base_size) / interpolation_scale
    gw = np.tile(pos[None, :], (grid, 1)).reshape(-1)
    gh = np.tile(pos[:, None], (1, grid)).reshape(-1)
    quarter = dim // 4
    omega = 1.0 / 10000 ** (np.arange(quarter, dtype=np.float64) / quarter)

    def one(p):
        o = p[:, None] * omega[None]
        return np.concatenate([np.sin(o), np.cos(o)], axis=1)
    return torch.from_numpy(np.concatenate([one(gw), one(gh)], axis=1)).float()


def unet_param_specs(cfg, layers_per_block=2):
    """(name, shape) of every UNet parameter, diffusers state_dict naming (unet_2d_condition.py:171-484)."""
    bo = cfg["block_out"]
    temb = bo[0] * 4
    out = []

    def lin(n, o, i, bias=True):
        out.append((n + ".weight", (o, i)))
        if bias:
            out.append((n + ".bias", (o,)))

    def conv(n, o, i, k):
        out.append((n + ".weight", (o, i, k, k)))
        out.append((n + ".bias", (o,)))

    def norm(n, c):
        out.append((n + ".weight", (c,)))
        out.append((n + ".bias", (c,)))

    def resnet(n, cin, cout, temb_ch):
        norm(n + ".norm1", cin)
        conv(n + ".conv1", cout, cin, 3)
        if temb_ch:
            lin(n + ".time_emb_proj", cout, temb_ch)
        norm(n + ".norm2", cout)
        conv(n + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(n + ".conv_shortcut", cout, cin, 1)

    def vit(n, c, depth):
        norm(n + ".norm", c)
        if cfg["linear_proj"]:
            lin(n + ".proj_in", c, c)
        else:
            conv(n + ".proj_in", c, c, 1)
        for k in range(depth):
            b = "%s.transformer_blocks.%d" % (n, k)
            norm(b + ".norm1", c)
            for a, cd in (("attn1", c), ("attn2", cfg["ctx_dim"])):
                lin(b + "." + a + ".to_q", c, c, bias=False)
                lin(b + "." + a + ".to_k", c, cd, bias=False)
                lin(b + "." + a + ".to_v", c, cd, bias=False)
                lin(b + "." + a + ".to_out.0", c, c)
                if a == "attn1":
                    norm(b + ".norm2", c)
            norm(b + ".norm3", c)
            lin(b + ".ff.net.0.proj", 8 * c, c)
            lin(b + ".ff.net.2", c, 4 * c)
        if cfg["linear_proj"]:
            lin(n + ".proj_out", c, c)
        else:
            conv(n + ".proj_out", c, c, 1)

    conv("conv_in", bo[0], 4, 3)
    lin("time_embedding.linear_1", temb, bo[0])
    lin("time_embedding.linear_2", temb, temb)
    if cfg["add_time_dim"]:
        lin("add_embedding.linear_1", temb, cfg["add_in"])
        lin("add_embedding.linear_2", temb, temb)
    n = len(bo)
    ch = bo[0]
    skips = [bo[0]]
    for i in range(n):
        cin, ch = ch, bo[i]
        for j in range(layers_per_block):
            resnet("down_blocks.%d.resnets.%d" % (i, j), cin if j == 0 else ch, ch, temb)
            if cfg["down_attn"][i]:
                vit("down_blocks.%d.attentions.%d" % (i, j), ch, cfg["depth"][i])
            skips.append(ch)
        if i != n - 1:
            conv("down_blocks.%d.downsamplers.0.conv" % i, ch, ch, 3)
            skips.append(ch)
    resnet("mid_block.resnets.0", ch, ch, temb)
    vit("mid_block.attentions.0", ch, cfg["depth"][-1])
    resnet("mid_block.resnets.1", ch, ch, temb)
    prev = bo[-1]
    for i in range(n):
        li = n - 1 - i
        cout = bo[li]
        for j in range(layers_per_block + 1):
            sk = skips.pop()
            resnet("up_blocks.%d.resnets.%d" % (i, j), (prev if j == 0 else cout) + sk, cout, temb)
            if cfg["up_attn"][i]:
                vit("up_blocks.%d.attentions.%d" % (i, j), cout, cfg["depth"][li])
        if i != n - 1:
            conv("up_blocks.%d.upsamplers.0.conv" % i, cout, cout, 3)
        prev = cout
    norm("conv_norm_out", bo[0])
    conv("conv_out", 4, bo[0], 3)
    return out


def vae_param_specs(cfg):
    """(name, shape) of the VAE-encoder parameters ('encoder.*', 'quant_conv.*'), diffusers naming."""
    bo = cfg["block_out"]
    out = []

    def conv(n, o, i, k):
        out.append((n + ".weight", (o, i, k, k)))
        out.append((n + ".bias", (o,)))

    def norm(n, c):
        out.append((n + ".weight", (c,)))
        out.append((n + ".bias", (c,)))

    def lin(n, o, i):
        out.append((n + ".weight", (o, i)))
        out.append((n + ".bias", (o,)))

    def resnet(n, cin, cout):
        norm(n + ".norm1", cin)
        conv(n + ".conv1", cout, cin, 3)
        norm(n + ".norm2", cout)
        conv(n + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(n + ".conv_shortcut", cout, cin, 1)

    conv("encoder.conv_in", bo[0], 3, 3)
    ch = bo[0]
    for i, co in enumerate(bo):
        for j in range(cfg["layers"]):
            resnet("encoder.down_blocks.%d.resnets.%d" % (i, j), ch if j == 0 else co, co)
        ch = co
        if i != len(bo) - 1:
            conv("encoder.down_blocks.%d.downsamplers.0.conv" % i, ch, ch, 3)
    resnet("encoder.mid_block.resnets.0", ch, ch)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        lin(a + "." + p, ch, ch)
    resnet("encoder.mid_block.resnets.1", ch, ch)
    norm("encoder.conv_norm_out", ch)
    conv("encoder.conv_out", 2 * cfg["latent"], ch, 3)
    if cfg.get("quant_conv", True):
        conv("quant_conv", 2 * cfg["latent"], 2 * cfg["latent"], 1)
    return out


def vae_decoder_param_specs(cfg):
    """(name, shape) of the VAE-DECODER parameters ('decoder.*', 'post_quant_conv.*'), diffusers naming [autoencoders/vae.py
    Decoder + AutoencoderKL.post_quant_conv, un-vendored]. Optional: only the `vae-out` path (diffusion_feature.py:477-485)
    reads them; a pipe without them raises when 'vae-out' is requested."""
    bo = list(reversed(cfg["block_out"]))
    out = []

    def conv(n, o, i, k):
        out.append((n + ".weight", (o, i, k, k)))
        out.append((n + ".bias", (o,)))

    def norm(n, c):
        out.append((n + ".weight", (c,)))
        out.append((n + ".bias", (c,)))

    def resnet(n, cin, cout):
        norm(n + ".norm1", cin)
        conv(n + ".conv1", cout, cin, 3)
        norm(n + ".norm2", cout)
        conv(n + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(n + ".conv_shortcut", cout, cin, 1)

    if cfg.get("quant_conv", True):
        conv("post_quant_conv", cfg["latent"], cfg["latent"], 1)
    ch = bo[0]
    conv("decoder.conv_in", ch, cfg["latent"], 3)
    resnet("decoder.mid_block.resnets.0", ch, ch)
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", ch)
    for p in ("to_q", "to_k", "to_v", "to_out.0"):
        out.append((a + "." + p + ".weight", (ch, ch)))
        out.append((a + "." + p + ".bias", (ch,)))
    resnet("decoder.mid_block.resnets.1", ch, ch)
    for i, co in enumerate(bo):
        for j in range(cfg["layers"] + 1):
            resnet("decoder.up_blocks.%d.resnets.%d" % (i, j), ch if j == 0 else co, co)
        ch = co
        if i != len(bo) - 1:
            conv("decoder.up_blocks.%d.upsamplers.0.conv" % i, ch, ch, 3)
    norm("decoder.conv_norm_out", ch)
    conv("decoder.conv_out", 3, ch, 3)
    return out


_RESIDUAL_OUT = ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight")


def init_param(name, shape, device="cpu"):
    """Deterministic synthetic weight for a parameter name: seed = crc32(name);
    weights ~ N(0, 1/fan_in) (x0.3 on residual-branch output layers so that deep random stacks stay
    well conditioned), norm gamma = 1 + 0.02 N(0,1), norm beta / biases = 0.02 N(0,1).
    device='cpu' is bit-reproducible everywhere (used by tests and golden fixtures); a CUDA device uses the
    CUDA generator (fast path for the 2.6 B-parameter SDXL bench; the oracle then takes .cpu() copies)."""
    g = torch.Generator(device=device)
    g.manual_seed(zlib.crc32(name.encode()))
    r = torch.randn(*shape, generator=g, device=device, dtype=torch.float32)
    is_norm = (".norm" in name or "group_norm" in name or "conv_norm_out" in name) and ".linear." not in name
    if name.endswith(".bias"):
        return 0.02 * r
    if is_norm:
        return 1.0 + 0.02 * r
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 0.3 if name.endswith(_RESIDUAL_OUT) else 1.0
    return r * (gain / fan_in ** 0.5)


def synthetic_state_dict(version, device="cpu", unet_cfg=None, vae_cfg=None, dit_cfg=None, flux_cfg=None,
                         with_decoder=False):
    """name -> fp32 tensor for 'unet.*' (or 'transformer.*') and 'vae.*' (random init; no checkpoints offline).
    with_decoder: also the VAE decoder ('vae.decoder.*', 'vae.post_quant_conv.*') that only `vae-out` needs."""
    vcfg = vae_cfg or VAE_CONFIGS[version]
    sd = {}
    if with_decoder:
        for n, s_ in vae_decoder_param_specs(vcfg):
            sd["vae." + n] = init_param("vae." + n, s_, device)
    fcfg = flux_cfg or (FLUX_CONFIGS.get(version) if (unet_cfg is None and dit_cfg is None) else None)
    if fcfg is not None:
        for n, s in flux_param_specs(fcfg):
            sd["transformer." + n] = init_param("transformer." + n, s, device)
        for n, s in vae_param_specs(vcfg):
            sd["vae." + n] = init_param("vae." + n, s, device)
        return sd
    dcfg = dit_cfg or (DIT_CONFIGS.get(version) if unet_cfg is None else None)
    if dcfg is not None:
        C = dcfg["heads"] * dcfg["head_dim"]
        grid = dcfg["sample_size"] // dcfg["patch"]
        for n, s in dit_param_specs(dcfg):
            key = "transformer." + n
            if n == "pos_embed.pos_embed":
                sd[key] = sincos_pos_embed_2d(C, grid, grid, dcfg["interpolation_scale"])[None].to(device)
            elif n.endswith("scale_shift_table"):
                g = torch.Generator(device=device)
                g.manual_seed(zlib.crc32(key.encode()))
                sd[key] = torch.randn(*s, generator=g, device=device) / C ** 0.5   # diffusers init: randn / sqrt(dim)
            else:
                sd[key] = init_param(key, s, device)
        for n, s in vae_param_specs(vcfg):
            sd["vae." + n] = init_param("vae." + n, s, device)
        return sd
    ucfg = unet_cfg or UNET_CONFIGS[version]
    for n, s in unet_param_specs(ucfg):
        sd["unet." + n] = init_param("unet." + n, s, device)
    for n, s in vae_param_specs(vcfg):
        sd["vae." + n] = init_param("vae." + n, s, device)
    return sd


def _unet_arch(cfg):
    a = UNetArch()
    a.in_channels, a.out_channels = 4, 4
    a.num_levels = len(cfg["block_out"])
    a.layers_per_block = 2
    for i, v in enumerate(cfg["block_out"]):
        a.block_out_channels[i] = v
        a.down_has_attn[i] = cfg["down_attn"][i]
        a.up_has_attn[i] = cfg["up_attn"][i]
        a.transformer_depth[i] = cfg["depth"][i]
        a.num_heads[i] = cfg["heads"][i]
    a.cross_attention_dim = cfg["ctx_dim"]
    a.use_linear_projection = int(cfg["linear_proj"])
    a.addition_time_embed_dim = cfg["add_time_dim"]
    a.projection_class_embeddings_input_dim = cfg["add_in"]
    a.norm_num_groups = 32
    a.norm_eps = cfg["eps"]
    return a


def _vae_arch(cfg):
    a = VaeArch()
    a.in_channels, a.latent_channels = 3, cfg["latent"]
    a.num_levels = len(cfg["block_out"])
    for i, v in enumerate(cfg["block_out"]):
        a.block_out_channels[i] = v
    a.layers_per_block = cfg["layers"]
    a.norm_num_groups = 32
    a.norm_eps = cfg["eps"]
    a.scaling_factor = cfg["scaling_factor"]
    a.shift_factor = cfg.get("shift_factor", 0.0)
    return a


def _dit_arch(cfg):
    a = DitArch()
    a.in_channels, a.out_channels, a.patch_size = cfg["in_ch"], cfg["out_ch"], cfg["patch"]
    a.num_layers, a.num_heads, a.head_dim = cfg["layers"], cfg["heads"], cfg["head_dim"]
    a.caption_channels = cfg["caption_dim"]
    a.norm_eps = cfg["eps"]
    return a


def _flux_arch(cfg):
    a = FluxArch()
    a.in_channels, a.num_layers, a.num_single_layers = cfg["in_ch"], cfg["layers"], cfg["single_layers"]
    a.num_heads, a.head_dim = cfg["heads"], cfg["head_dim"]
    a.joint_attention_dim, a.pooled_projection_dim = cfg["joint_dim"], cfg["pooled_dim"]
    a.guidance_embeds = int(cfg["guidance_embeds"])
    return a


def expected_shapes(unet_cfg=None, vae_cfg=None, dit_cfg=None, flux_cfg=None):
    """name -> shape of every parameter the architecture reads (prefixed 'unet.' / 'transformer.' / 'vae.')."""
    out = {}
    if flux_cfg is not None:
        out.update({"transformer." + n: tuple(s) for n, s in flux_param_specs(flux_cfg)})
    elif dit_cfg is not None:
        out.update({"transformer." + n: tuple(s) for n, s in dit_param_specs(dit_cfg)})
    elif unet_cfg is not None:
        out.update({"unet." + n: tuple(s) for n, s in unet_param_specs(unet_cfg)})
    if vae_cfg is not None:
        out.update({"vae." + n: tuple(s) for n, s in vae_param_specs(vae_cfg)})
    return out


def optional_shapes(vae_cfg):
    """name -> shape of the parameters a checkpoint MAY bring: the VAE decoder of the `vae-out` path."""
    return {"vae." + n: tuple(s) for n, s in vae_decoder_param_specs(vae_cfg)} if vae_cfg is not None else {}


# AutoencoderKL checkpoints written before diffusers 0.18 name the mid-block attention projections query / key /
# value / proj_attn ([diffusers modeling_utils._convert_deprecated_attention_blocks]); some store them as 1x1 convs
_VAE_ATTN_RENAMES = {".query.": ".to_q.", ".key.": ".to_k.", ".value.": ".to_v.", ".proj_attn.": ".to_out.0."}


def _read_safetensors_dir(folder):
    """Every tensor of `folder`/*.safetensors (single file or shards), as CPU tensors."""
    import glob
    import os
    from safetensors import safe_open
    files = sorted(glob.glob(os.path.join(folder, "*.safetensors")))
    if not files:
        raise FileNotFoundError("no .safetensors file in %s" % folder)
    # prefer the full-precision file when fp16 variants sit next to it (diffusion_pytorch_model[.fp16].safetensors)
    plain = [f for f in files if ".fp16." not in os.path.basename(f)]
    out = {}
    for f in (plain or files):
        with safe_open(f, framework="pt", device="cpu") as sf:
            for k in sf.keys():
                out[k] = sf.get_tensor(k)
    return out


def load_diffusers_dir(model_dir, version, with_decoder=False):
    """Read a diffusers-layout checkpoint directory (the layout `from_pretrained(...).save_pretrained(dir)` writes and
    the hub snapshots models.py:18-172 downloads): <dir>/unet/*.safetensors (or <dir>/transformer/ for the DiT / Flux
    families) and <dir>/vae/*.safetensors. Returns the name -> tensor dict `B200Pipe.load_state_dict` takes: parameter
    names exactly as diffusers writes them, prefixed 'unet.' / 'transformer.' / 'vae.' (decoder tensors are dropped
    unless with_decoder: only `vae-out` decodes)."""
    import os
    if not os.path.isdir(model_dir):
        raise FileNotFoundError("model directory %s does not exist" % model_dir)
    is_tr = version in DIT_CONFIGS or version in FLUX_CONFIGS
    sub = "transformer" if is_tr else "unet"
    sd = {}
    for k, v in _read_safetensors_dir(os.path.join(model_dir, sub)).items():
        sd[sub + "." + k] = v
    for k, v in _read_safetensors_dir(os.path.join(model_dir, "vae")).items():
        if not (k.startswith("encoder.") or k.startswith("quant_conv.") or
                (with_decoder and (k.startswith("decoder.") or k.startswith("post_quant_conv.")))):
            continue
        for a, b in _VAE_ATTN_RENAMES.items():
            k = k.replace(a, b)
        if ".attentions." in k and k.endswith(".weight") and v.dim() == 4 and v.shape[-2:] == (1, 1):
            v = v[:, :, 0, 0]
        sd["vae." + k] = v
    return sd


def save_diffusers_dir(sd, model_dir):
    """Inverse of load_diffusers_dir for a name -> tensor dict (used to stage synthetic checkpoints on disk)."""
    import os
    from safetensors.torch import save_file
    groups = {}
    for k, v in sd.items():
        sub, name = k.split(".", 1)
        groups.setdefault(sub, {})[name] = v.detach().cpu().contiguous()
    for sub, tensors in groups.items():
        os.makedirs(os.path.join(model_dir, sub), exist_ok=True)
        save_file(tensors, os.path.join(model_dir, sub, "diffusion_pytorch_model.safetensors"))


class B200Pipe:
    """What `FeatureExtractor` holds in place of a diffusers pipeline on the B200 path."""

    def __init__(self, version, unet_cfg, vae_cfg, device, dit_cfg=None, flux_cfg=None):
        self.version = version
        self.unet_cfg = unet_cfg
        self.dit_cfg = dit_cfg
        self.flux_cfg = flux_cfg
        self.vae_cfg = vae_cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GdfError("the B200 extraction path needs a CUDA device (got %s); there is no CPU fallback"
                                % device)
        self.lib = _lib.load()
        self.handle = ctypes.c_void_p()
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        va = _vae_arch(vae_cfg)
        if flux_cfg is not None:
            fa = _flux_arch(flux_cfg)
            check(self.lib.gdf_create_flux(ctypes.byref(fa), ctypes.byref(va), self.dev_index,
                                           ctypes.byref(self.handle)))
        elif dit_cfg is not None:
            da = _dit_arch(dit_cfg)
            check(self.lib.gdf_create_dit(ctypes.byref(da), ctypes.byref(va), self.dev_index,
                                          ctypes.byref(self.handle)))
        else:
            ua = _unet_arch(unet_cfg)
            check(self.lib.gdf_create(ctypes.byref(ua), ctypes.byref(va), self.dev_index, ctypes.byref(self.handle)))
        self._finalized = False
        self.has_decoder = False          # set by load_state_dict when the checkpoint brings 'vae.decoder.*'

    def load_state_dict(self, sd, chunk=256):
        """sd: name -> tensor ('unet.*', 'vae.*'), any float dtype / device; uploaded as fp32. Every tensor the
        architecture reads is checked against its expected shape first (a checkpoint / config mismatch raises here
        instead of reaching the device); names the architecture does not know (decoder, EMA copies, ...) are skipped."""
        want = expected_shapes(self.unet_cfg, self.vae_cfg, self.dit_cfg, self.flux_cfg)
        opt = optional_shapes(self.vae_cfg)
        if any(n in sd for n in opt):        # a decoder comes as a whole or not at all
            missing = [n for n in opt if n not in sd]
            if missing:
                raise _lib.GdfError("gdf error -5: the checkpoint holds part of the VAE decoder only; missing %s"
                                    % ", ".join(missing[:6]))
            want = dict(want, **opt)
            self.has_decoder = True
        bad = ["%s: checkpoint %s, architecture %s" % (n, tuple(sd[n].shape), want[n])
               for n in want if n in sd and tuple(sd[n].shape) != tuple(want[n])
               and not (n.endswith("pos_embed.pos_embed"))]      # position table: validated against img_size by gdf_plan
        if bad:
            raise _lib.GdfError("gdf error -5: weight shapes do not match the architecture of '%s':\n  %s"
                                % (self.version, "\n  ".join(bad[:8]) + ("\n  ..." if len(bad) > 8 else "")))
        sd = {n: t for n, t in sd.items() if n in want or n.startswith("vae.quant_conv.")}
        if not self.vae_cfg.get("quant_conv", True) and "vae.quant_conv.weight" not in sd:
            # Flux VAE (use_quant_conv False): the executor folds conv_out . quant_conv, so feed it the identity
            nm = 2 * self.vae_cfg["latent"]
            sd = dict(sd)
            sd["vae.quant_conv.weight"] = torch.eye(nm).reshape(nm, nm, 1, 1)
            sd["vae.quant_conv.bias"] = torch.zeros(nm)
        names = list(sd.keys())
        with torch.cuda.device(self.dev_index):
            for s in range(0, len(names), chunk):
                part = names[s:s + chunk]
                tens = [sd[n].detach().to(self.device, torch.float32).contiguous() for n in part]
                c_names = (ctypes.c_char_p * len(part))(*[n.encode() for n in part])
                c_ptrs = (ctypes.c_void_p * len(part))(*[t.data_ptr() for t in tens])
                shapes = [d for t in tens for d in t.shape]
                c_shapes = (ctypes.c_int64 * len(shapes))(*shapes)
                c_ranks = (ctypes.c_int * len(part))(*[t.dim() for t in tens])
                check(self.lib.gdf_load_weights(self.handle, c_names, c_ptrs, c_shapes, c_ranks, len(part),
                                                _lib.stream_ptr()))
                del tens
        self._finalized = False

    def finalize(self):
        if not self._finalized:
            with torch.cuda.device(self.dev_index):
                check(self.lib.gdf_finalize_weights(self.handle, _lib.stream_ptr()))
            self._finalized = True

    def close(self):
        if self.handle:
            self.lib.gdf_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_diffusion_model(version, dtype, offline_lora=None, offline_lora_filename=None, device="cuda",
                        state_dict=None, unet_cfg=None, vae_cfg=None, weight_device=None, dit_cfg=None, flux_cfg=None,
                        model_dir=None, synthetic=False, with_decoder=False):
    """Mirror of feature/components/models.py:10 for the B200 path.

    Where the reference calls `Pipeline.from_pretrained(model_id)` (hub download, models.py:18-172), the weights come
    from, in this order:
      1. `state_dict`  - name -> tensor with diffusers parameter names prefixed 'unet.' / 'transformer.' / 'vae.';
      2. `model_dir`   - a diffusers-layout directory (<dir>/unet|transformer/*.safetensors + <dir>/vae/*.safetensors),
                         or the environment variable GDF_MODEL_DIR (the directory itself, or GDF_MODEL_DIR/<version>);
      3. `synthetic=True` (or GDF_SYNTHETIC=1) - deterministic random weights generated by parameter name
                         (`synthetic_state_dict`): benchmarks and parity tests, never a silent default.
    With none of the three the call raises: features of a random network are not what a caller of the reference's
    factory expects to get. with_decoder: also load (or generate) the VAE decoder, which only `vae-out` needs."""
    import os
    if dtype not in ("float32", "float16"):
        raise NotImplementedError                      # models.py:11-16
    if offline_lora is not None:
        raise NotImplementedError("LoRA loading is outside the B200 hot path (SURVEY.md 2.1 OUT OF SCOPE)")
    if version in _NOT_BUILT:
        raise NotImplementedError("version '%s' is not built on the B200 path (SURVEY.md 8f)" % version)
    is_flux = flux_cfg is not None or (version in FLUX_CONFIGS and unet_cfg is None and dit_cfg is None)
    is_dit = not is_flux and (dit_cfg is not None or (version in DIT_CONFIGS and unet_cfg is None))
    if not is_flux and not is_dit and version not in UNET_CONFIGS and unet_cfg is None:
        raise NotImplementedError                      # models.py:173-174
    fcfg = (flux_cfg or FLUX_CONFIGS[version]) if is_flux else None
    dcfg = (dit_cfg or DIT_CONFIGS[version]) if is_dit else None
    ucfg = None if (is_flux or is_dit) else (unet_cfg or UNET_CONFIGS[version])
    vcfg = vae_cfg or VAE_CONFIGS[version]
    if state_dict is None:
        env_dir = os.environ.get("GDF_MODEL_DIR")
        if model_dir is None and env_dir:
            model_dir = os.path.join(env_dir, version) if os.path.isdir(os.path.join(env_dir, version)) else env_dir
        if model_dir is not None:
            state_dict = load_diffusers_dir(model_dir, version, with_decoder=with_decoder)
        elif synthetic or os.environ.get("GDF_SYNTHETIC") == "1":
            state_dict = synthetic_state_dict(version, weight_device or "cpu", ucfg, vcfg, dcfg, fcfg,
                                              with_decoder=with_decoder)
        else:
            raise _lib.GdfError(
                "get_diffusion_model('%s'): no weights. Pass state_dict=, model_dir= (or set GDF_MODEL_DIR) pointing at "
                "a diffusers-layout checkpoint, or ask for random weights explicitly with synthetic=True / "
                "GDF_SYNTHETIC=1 (there is no network for from_pretrained)" % version)
    pipe = B200Pipe(version, ucfg, vcfg, device, dit_cfg=dcfg, flux_cfg=fcfg)
    pipe.load_state_dict(state_dict)
    pipe.finalize()
    return pipe
