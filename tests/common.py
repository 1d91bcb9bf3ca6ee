"""Shared helpers of the test-suite: seeded synthetic inputs and oracle construction."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import torch_oracle as O  # noqa: E402  (test infrastructure: the checker)

# reduced-width configurations with the SDXL / SD-2.1 topology (every kernel shape class of the full models:
# skip-concat resnets with 1x1 shortcuts, stride-2 and upsampled convs, multi-depth transformers, head_dim 64)
TINY_XL = dict(block_out=(64, 128, 256), down_attn=(0, 1, 1), up_attn=(1, 1, 0), depth=(1, 1, 2),
               heads=(1, 2, 4), ctx_dim=128, linear_proj=True, add_time_dim=32, add_in=64 + 6 * 32, eps=1e-5)
TINY_21 = dict(block_out=(64, 128, 256, 256), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
               heads=(1, 2, 4, 4), ctx_dim=128, linear_proj=True, add_time_dim=0, add_in=0, eps=1e-5)
# SD-1.5 topology: 1x1-conv proj_in/out, 8 heads per level -> head dims 8 / 16 / 32 / 32 (generic attention kernel)
TINY_15 = dict(block_out=(64, 128, 256, 256), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
               heads=(8, 8, 8, 8), ctx_dim=128, linear_proj=False, add_time_dim=0, add_in=0, eps=1e-5)
# PixArt-Sigma topology at reduced size: the real head_dim (72 -> padded-head-dim attention kernel), hidden 576,
# patch 2 on a 16x16 latent (128x128 images) -> 64 tokens, caption width 256
TINY_DIT = dict(layers=2, heads=8, head_dim=72, in_ch=4, out_ch=8, patch=2, caption_dim=256, sample_size=16,
                interpolation_scale=2.0, eps=1e-6)
TINY_VAE = dict(block_out=(64, 64, 128, 128), layers=2, latent=4, eps=1e-6, scaling_factor=0.13025)
# Flux topology at reduced size with the real head_dim and rotary axes: 2 double + 2 single MMDiT blocks, hidden 256
# (2 heads x 128), 16 latent channels packed 2x2 -> 64 input channels, 16 text tokens of width 64, pooled width 32;
# 128x128 images -> 8x8 = 64 image tokens
TINY_FLUX = dict(layers=2, single_layers=2, heads=2, head_dim=128, in_ch=64, joint_dim=64, pooled_dim=32,
                 guidance_embeds=True, axes_dims_rope=(16, 56, 56), ctx_len=16)
TINY_VAE_FLUX = dict(block_out=(64, 64, 128, 128), layers=2, latent=16, eps=1e-6, scaling_factor=0.3611,
                     shift_factor=0.1159, quant_conv=False)


def make_inputs(batch, img, ctx_dim, pooled_dim=None, ctx_len=77):
    """Synthetic inputs of SURVEY.md 8d: images rand*2-1 (seed 1234), ctx N(0,1) (1235), pooled (1236),
    eps_vae / eps_q (1237 / 1238)."""
    g = lambda s: torch.Generator().manual_seed(s)
    image = torch.rand(batch, 3, img, img, generator=g(1234)) * 2 - 1
    ctx = torch.randn(1, ctx_len, ctx_dim, generator=g(1235))
    pooled = torch.randn(1, pooled_dim, generator=g(1236)) if pooled_dim else None
    L = img // 8
    eps_vae = torch.randn(batch, 4, L, L, generator=g(1237))
    eps_q = torch.randn(batch, 4, L, L, generator=g(1238))
    return image, ctx, pooled, eps_vae, eps_q


def build_oracle(unet_cfg, vae_cfg, sd):
    """Oracle modules loaded from a 'unet.*' / 'vae.*' state dict (fp32, CPU)."""
    unet = O.UNet2DConditionModel(unet_cfg)
    vae = O.Vae(vae_cfg["scaling_factor"], decoder=any(k.startswith("vae.decoder.") for k in sd),
                block_out=vae_cfg["block_out"], layers=vae_cfg["layers"], latent=vae_cfg["latent"], eps=vae_cfg["eps"])
    usd = {k[len("unet."):]: v.float().cpu() for k, v in sd.items() if k.startswith("unet.")}
    vsd = {k[len("vae."):]: v.float().cpu() for k, v in sd.items() if k.startswith("vae.")}
    unet.load_state_dict(usd, strict=True)
    vae.load_state_dict(vsd, strict=True)
    return unet.eval(), vae.eval()


def build_oracle_dit(dit_cfg, vae_cfg, sd):
    """Oracle PixArt transformer + VAE from a 'transformer.*' / 'vae.*' state dict (fp32, CPU)."""
    model = O.PixArtTransformer2DModel(dit_cfg)
    vae = O.Vae(vae_cfg["scaling_factor"], block_out=vae_cfg["block_out"], layers=vae_cfg["layers"],
                latent=vae_cfg["latent"], eps=vae_cfg["eps"])
    tsd = {k[len("transformer."):]: v.float().cpu() for k, v in sd.items() if k.startswith("transformer.")}
    vsd = {k[len("vae."):]: v.float().cpu() for k, v in sd.items() if k.startswith("vae.")}
    model.load_state_dict(tsd, strict=True)
    vae.load_state_dict(vsd, strict=True)
    return model.eval(), vae.eval()


def build_oracle_flux(flux_cfg, vae_cfg, sd):
    """Oracle Flux transformer + VAE from a 'transformer.*' / 'vae.*' state dict (fp32, CPU)."""
    model = O.FluxTransformer2DModel(flux_cfg)
    vae = O.Vae(vae_cfg["scaling_factor"], vae_cfg.get("shift_factor", 0.0), vae_cfg.get("quant_conv", True),
                block_out=vae_cfg["block_out"], layers=vae_cfg["layers"], latent=vae_cfg["latent"], eps=vae_cfg["eps"])
    tsd = {k[len("transformer."):]: v.float().cpu() for k, v in sd.items() if k.startswith("transformer.")}
    vsd = {k[len("vae."):]: v.float().cpu() for k, v in sd.items() if k.startswith("vae.")}
    model.load_state_dict(tsd, strict=True)
    vae.load_state_dict(vsd, strict=True)
    return model.eval(), vae.eval()


def make_flux_inputs(batch, img, flux_cfg, latent=16):
    """Images as make_inputs; T5-like context (1, ctx_len, joint_dim) seed 1235, pooled (1, pooled_dim) seed 1236,
    eps_vae / eps_q with `latent` channels (seeds 1237 / 1238)."""
    g = lambda s: torch.Generator().manual_seed(s)
    image = torch.rand(batch, 3, img, img, generator=g(1234)) * 2 - 1
    ctx = torch.randn(1, flux_cfg["ctx_len"], flux_cfg["joint_dim"], generator=g(1235))
    pooled = torch.randn(1, flux_cfg["pooled_dim"], generator=g(1236))
    L = img // 8
    eps_vae = torch.randn(batch, latent, L, L, generator=g(1237))
    eps_q = torch.randn(batch, latent, L, L, generator=g(1238))
    return image, ctx, pooled, eps_vae, eps_q


def make_dit_inputs(batch, img, caption_dim, ctx_len=24, masked_tail=5):
    """Images / noise as make_inputs; caption embeddings N(0,1) seed 1235 (1, ctx_len, caption_dim) and an
    attention mask whose last `masked_tail` tokens are padding (0), like a tokenised short prompt."""
    image, ctx, _, eps_vae, eps_q = make_inputs(batch, img, caption_dim, None, ctx_len)
    mask = torch.ones(1, ctx_len)
    if masked_tail:
        mask[:, ctx_len - masked_tail:] = 0
    return image, ctx, mask, eps_vae, eps_q


def compare_maps(got, want):
    """Per-map cosine similarity / relative L2 / max-relative error between two dicts of (B,C,h,w) maps."""
    rows = []
    for k, w in want.items():
        g = got[k].float().cpu()
        w = w.float()
        assert g.shape == w.shape, "%s: shape %s vs %s" % (k, tuple(g.shape), tuple(w.shape))
        gf, wf = g.flatten(), w.flatten()
        cos = torch.nn.functional.cosine_similarity(gf, wf, dim=0).item()
        rel = ((gf - wf).norm() / (wf.norm() + 1e-12)).item()
        maxrel = ((gf - wf).abs().max() / (wf.abs().max() + 1e-12)).item()
        rows.append((k, cos, rel, maxrel))
    return rows


# ------------------------------------------------------------------------------------------ CLI / on-disk format helpers
# (shared by tools/make_golden.py, which runs the reference's own extract_feature.py on them, and tests/test_cpu.py)
import numpy as np  # noqa: E402


def cli_fixture_inputs(root):
    """Six solid-colour PNGs in two folders (red channel = image index) + a prompt file; shared with tests/test_cpu.py."""
    from PIL import Image
    paths = []
    for i in range(6):
        d = os.path.join(root, "imgs", "cat" if i < 3 else "dog")
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, "im%d.png" % i)
        Image.new("RGB", (40, 24), (10 * i + 5, 0, 0)).save(p)
        paths.append(p)
    with open(os.path.join(root, "prompt.txt"), "w") as f:
        f.write("a photo")
    return paths


class FakeExtractor:
    """Stands in for FeatureExtractor under the CLI: maps whose values depend only on the image (red channel)."""
    LAYERS = (("up-level1-repeat1-vit-block0-self-q", 6, 8), ("mid-vit-out", 4, 4), ("unet-out", 2, 16))

    def __init__(self, *a, **k):
        self.feature_store = type("S", (), {"accept_all": False})()

    def encode_prompt(self, p):
        return p

    def extract(self, prompts, n, images, t=None, **kw):
        out = {}
        for name, c, hw in self.LAYERS:
            rows = []
            for im in images:
                idx = im.convert("RGB").getpixel((0, 0))[0]
                g = torch.Generator().manual_seed(1000 + idx)
                rows.append(torch.randn(c, hw, hw, generator=g))
            out[name] = torch.stack(rows).to(torch.float16)
        return out


CLI_MODES = {
    "default": ["--split", "val"],
    "sample_first_original": ["--sample_name_first", "--use_original_filename"],
    "aggregate_nested": ["--aggregate_output", "--use_original_filename", "--nested_input_dir"],
}


def tree_digest(root):
    import hashlib
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            p = os.path.join(d, f)
            a = np.load(p)
            out[os.path.relpath(p, root)] = {"shape": list(a.shape), "dtype": str(a.dtype),
                                             "sha1": hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()}
    return out


