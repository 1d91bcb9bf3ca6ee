"""Generate the golden fixtures under tests/golden/ by EXECUTING the reference's own code in this container.

    python tools/make_golden.py            (needs /root/reference; cannot run on the GPU box)

Fixtures (small, committed; regenerate with this script):
  unet_tiny_xl.pt / unet_tiny_21.pt
      Output of the reference's vendored UNet2DConditionModel.forward (feature/diffusers/models/unet/
      unet_2d_condition.py) + the reference's real prepare_feature_extractor / FeatureStore
      (feature/components/feature_extractor.py, train_unet=True so nothing is cast) on the reduced-width
      topologies of tests/common.py, synthetic weights by name, seeded inputs. Stored: every feature map (fp16),
      the noise prediction, and the inputs. The oracle and the CUDA path are both checked against these.
  flux_tiny.pt
      Output of the reference's vendored FluxTransformer2DModel (feature/diffusers/models/transformers/
      transformer_flux.py: ctor, forward, both block classes) + vendored Attention / FluxAttnProcessor2_0 / FeedForward
      + the reference's real prepare_feature_extractor (Flux branch) / FeatureStore on the reduced topology
      TINY_FLUX; the un-vendored normalisation / embedding helpers come from tools/ref_shim.py.
  correspondence.pt
      Output of the reference's correspondence_utils.find_nn_source_correspondences / points_to_idxs
      (correspondence/correspondence/correspondence_utils.py:113-146) on seeded feature stacks and query points.
  extract_tiny_xl.pt
      Whole-path digest (VAE encode + q_sample + UNet) produced by the ORACLE (the VAE encoder and the schedulers are
      un-vendored in the reference, so this one is an oracle self-consistency fixture, not a reference pin).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_shim  # noqa: E402
from common import (CLI_MODES, FakeExtractor, cli_fixture_inputs, tree_digest,  # noqa: E402
                    O, TINY_15, TINY_21, TINY_DIT, TINY_FLUX, TINY_VAE, TINY_VAE_FLUX, TINY_XL, build_oracle,  # noqa: E402
                    build_oracle_dit, build_oracle_flux, make_dit_inputs, make_flux_inputs, make_inputs)
from generic_diffusion_feature_b200.components import models  # noqa: E402
from generic_diffusion_feature_b200.components.feature_extractor import (_dit_feature_ids, _flux_feature_ids,  # noqa: E402
                                                                        _unet_feature_ids)

OUT = os.path.join(ROOT, "tests", "golden")


def unet_inputs(cfg, B=1, L=8, seed=4321):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 4, L, L, generator=g)
    ctx = torch.randn(B, 77, cfg["ctx_dim"], generator=g)
    pooled = torch.randn(B, cfg["add_in"] - 6 * cfg["add_time_dim"], generator=g) if cfg["add_time_dim"] else None
    return x, ctx, pooled


def golden_unet(version, cfg, name):
    sd = models.synthetic_state_dict(version, "cpu", cfg, TINY_VAE)
    ref_unet = ref_shim.build_reference_unet(cfg)
    ref_unet.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("unet.")}, strict=True)
    ref_unet.eval()
    rfe = ref_shim.load_reference_feature_extractor()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.unet = ref_unet
    ids = _unet_feature_ids(cfg)
    store = rfe.prepare_feature_extractor(version, pipe, {i: True for i in ids}, 1, True)
    x, ctx, pooled = unet_inputs(cfg)
    kw, okw = {}, {}
    if pooled is not None:
        tid = O.add_time_ids(8 * x.shape[-1]).repeat(x.shape[0], 1)
        kw["added_cond_kwargs"] = {"text_embeds": pooled, "time_ids": tid}
        okw = dict(text_embeds=pooled, time_ids=tid)
    with torch.no_grad():
        out = ref_unet(x, timestep=torch.tensor([50.0]), encoder_hidden_states=ctx, return_dict=False, **kw)[0]
    feats = store.stored_feats
    assert list(feats.keys()) == ids
    # oracle check before anything is written
    ounet, _ = build_oracle(cfg, TINY_VAE, sd)
    ostore = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers(ounet, ostore)
    with torch.no_grad():
        oout = ounet(x, 50.0, ctx, **okw)
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in ids)
    print("%s: %d maps from the reference's vendored modules; oracle max |diff| %.2e (out %.2e)"
          % (name, len(ids), worst, (out - oout).abs().max().item()))
    assert worst < 1e-3
    torch.save({"version": version, "ids": ids, "x": x, "ctx": ctx, "pooled": pooled, "timestep": 50.0,
                "noise_pred": out, "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference vendored modules)"},
               os.path.join(OUT, name))


def golden_unet_maps(name="unet_tiny_xl_maps.pt"):
    """Attention-probability maps: the reference's REAL AttnStoreProcessor / AttentionStore /
    register_attention_store (feature/components/attention.py:102-263, 531-566) installed on the vendored UNet, the
    `...-self-map` / `...-cross-map` ids through the real FeatureStore, and the aggregated `attn` feature assembled
    like diffusion_feature.py:488-500. The oracle must reproduce all of it."""
    import torch.nn.functional as F
    version, cfg, img = "xl", TINY_XL, 64
    sd = models.synthetic_state_dict(version, "cpu", cfg, TINY_VAE)
    ref_unet = ref_shim.build_reference_unet(cfg)
    ref_unet.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("unet.")}, strict=True)
    ref_unet.eval()
    rfe = ref_shim.load_reference_feature_extractor()
    rat = ref_shim.load_reference_attention_store()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.unet = ref_unet
    base = _unet_feature_ids(cfg)
    map_ids = []
    for i in base:
        if i.endswith("-self-q"):
            map_ids.append(i[:-len("self-q")] + "self-map")
        if i.endswith("-cross-q"):
            map_ids.append(i[:-len("cross-q")] + "cross-map")
    layer = {i: True for i in base + map_ids}
    store = rfe.prepare_feature_extractor(version, pipe, layer, 1, True)
    categories = ["up_cross", "down_self", "mid_cross"]
    astore = rat.register_attention_store(version, pipe, img, True)
    x, ctx, pooled = unet_inputs(cfg, L=img // 8)
    tid = O.add_time_ids(img).repeat(x.shape[0], 1)
    with torch.no_grad():
        out = ref_unet(x, timestep=torch.tensor([50.0]), encoder_hidden_states=ctx, return_dict=False,
                       added_cond_kwargs={"text_embeds": pooled, "time_ids": tid})[0]
    feats = dict(store.stored_feats)
    got_maps = [k for k in feats if k.endswith("-map")]
    assert sorted(got_maps) == sorted(map_ids), (len(got_maps), len(map_ids))
    all_attns = []
    for category, maps in astore.aggregate_attention(categories).items():
        for size, attn in maps.items():
            all_attns.append(F.interpolate(attn, size=(img // 8, img // 8)))
    attn_feat = torch.cat(all_attns, dim=-3)
    # oracle
    ounet, _ = build_oracle(cfg, TINY_VAE, sd)
    ostore = O.FeatureStore(layer)
    O.attach_gatherers(ounet, ostore)
    oast = O.register_attention_store(ounet, img)
    with torch.no_grad():
        oout = ounet(x, 50.0, ctx, text_embeds=pooled, time_ids=tid)
    assert list(ostore.feats.keys()) == list(feats.keys())
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in feats)
    oattn = O.aggregated_attention_feature(oast, categories, img)
    print("%s: %d ids (%d maps) + aggregated attn %s from the reference's AttnStoreProcessor; oracle max |diff| %.2e "
          "(attn %.2e, out %.2e)" % (name, len(feats), len(map_ids), tuple(attn_feat.shape), worst,
                                     (attn_feat - oattn).abs().max().item(), (out - oout).abs().max().item()))
    assert worst < 1e-4 and (attn_feat - oattn).abs().max().item() < 1e-5
    torch.save({"ids": list(feats.keys()), "map_ids": map_ids, "categories": categories, "img": img, "x": x, "ctx": ctx,
                "pooled": pooled, "timestep": 50.0, "attn": attn_feat,
                "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference AttnStoreProcessor / AttentionStore)"},
               os.path.join(OUT, name))


def golden_unet_control(name="unet_tiny_xl_control.pt"):
    """ControlNet residual inputs: the reference's vendored UNet2DConditionModel.forward with
    down_block_additional_residuals / mid_block_additional_residual (unet_2d_condition.py:1236-1275) on seeded
    residual tensors; a handful of maps that depend on them + the noise prediction."""
    version, cfg = "xl", TINY_XL
    sd = models.synthetic_state_dict(version, "cpu", cfg, TINY_VAE)
    ref_unet = ref_shim.build_reference_unet(cfg)
    ref_unet.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("unet.")}, strict=True)
    ref_unet.eval()
    rfe = ref_shim.load_reference_feature_extractor()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.unet = ref_unet
    ids = [i for i in _unet_feature_ids(cfg) if i.startswith("up-") and i.endswith("-res-out")] + ["mid-vit-out", "unet-out"]
    order = [i for i in _unet_feature_ids(cfg) if i in set(ids)]
    store = rfe.prepare_feature_extractor(version, pipe, {i: True for i in order}, 1, True)
    x, ctx, pooled = unet_inputs(cfg)
    L = x.shape[-1]
    tid = O.add_time_ids(8 * L).repeat(x.shape[0], 1)
    # skip tensors of the oracle give the shapes: conv_in output + every resnet / downsampler output of the down path
    ounet, _ = build_oracle(cfg, TINY_VAE, sd)
    g = torch.Generator().manual_seed(99)
    bo = cfg["block_out"]
    shapes, res = [(bo[0], L)], L
    for i, c in enumerate(bo):
        shapes += [(c, res), (c, res)]
        if i != len(bo) - 1:
            res //= 2
            shapes.append((c, res))
    down = [0.3 * torch.randn(1, c, r, r, generator=g) for c, r in shapes]
    mid = 0.3 * torch.randn(1, bo[-1], res, res, generator=g)
    with torch.no_grad():
        out = ref_unet(x, timestep=torch.tensor([50.0]), encoder_hidden_states=ctx, return_dict=False,
                       added_cond_kwargs={"text_embeds": pooled, "time_ids": tid},
                       down_block_additional_residuals=tuple(down), mid_block_additional_residual=mid)[0]
        plain = ref_unet(x, timestep=torch.tensor([50.0]), encoder_hidden_states=ctx, return_dict=False,
                         added_cond_kwargs={"text_embeds": pooled, "time_ids": tid})[0]
    feats = dict(store.stored_feats)          # (second call overwrote nothing: reset between calls)
    ostore = O.FeatureStore({i: True for i in order})
    O.attach_gatherers(ounet, ostore)
    with torch.no_grad():
        oout = ounet(x, 50.0, ctx, text_embeds=pooled, time_ids=tid, down_residuals=down, mid_residual=mid)
    print("%s: noise prediction moves by %.3f with the residuals; oracle max |diff| %.2e"
          % (name, (out - plain).abs().max().item(), (out - oout).abs().max().item()))
    assert (out - oout).abs().max().item() < 1e-4 and (out - plain).abs().max().item() > 1e-2
    torch.save({"x": x, "ctx": ctx, "pooled": pooled, "timestep": 50.0, "down": down, "mid": mid, "noise_pred": out,
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference vendored UNet, ControlNet residuals)"},
               os.path.join(OUT, name))


def _build_ref_dit(cfg):
    """The reference's vendored BasicTransformerBlock stack inside the oracle's (un-vendored) outer PixArt model."""
    import torch.nn as nn
    root = ref_shim.install()
    C = cfg["heads"] * cfg["head_dim"]
    sd = models.synthetic_state_dict("pixart-sigma", "cpu", None, TINY_VAE, cfg)
    omodel, _ = build_oracle_dit(cfg, TINY_VAE, sd)

    class RefDit(nn.Module):
        def __init__(self):
            super().__init__()
            self.outer = omodel
            self.transformer_blocks = nn.ModuleList([
                root.attention.BasicTransformerBlock(C, cfg["heads"], cfg["head_dim"], cross_attention_dim=C,
                                                     activation_fn="gelu-approximate", attention_bias=True,
                                                     norm_type="ada_norm_single", norm_elementwise_affine=False,
                                                     norm_eps=cfg["eps"]) for _ in range(cfg["layers"])])

        def forward(self, sample, timestep, ctx, mask):
            o = self.outer
            B = sample.shape[0]
            bias = ((1 - mask) * -10000.0)[:, None, :]            # [pixart_transformer_2d.py forward head]
            x = o.pos_embed(sample)
            t6, emb = o.adaln_single(torch.full((B,), float(timestep)))
            c = o.caption_projection(ctx)
            for blk in self.transformer_blocks:
                x = blk(x, encoder_hidden_states=c, encoder_attention_mask=bias, timestep=t6)
            shift, scale = (o.scale_shift_table[None] + emb[:, None]).chunk(2, dim=1)
            x = o.proj_out(o.norm_out(x) * (1 + scale) + shift)
            g, p, oc = sample.shape[-1] // cfg["patch"], cfg["patch"], cfg["out_ch"]
            x = x.reshape(B, g, g, p, p, oc)
            return torch.einsum("nhwpqc->nchpwq", x).reshape(B, oc, g * p, g * p)

    ref = RefDit().eval()
    for i, blk in enumerate(ref.transformer_blocks):
        blk.load_state_dict({k[len("transformer.transformer_blocks.%d." % i):]: v for k, v in sd.items()
                             if k.startswith("transformer.transformer_blocks.%d." % i)}, strict=True)
    return ref, omodel, sd


def golden_dit_maps(name="dit_tiny_pixart_maps.pt"):
    """Attention-probability maps of the PixArt family: the reference's REAL AttnStoreProcessor /
    AttentionStore / register_attention_store (transformer branch, feature/components/attention.py:567-593: attn1 and
    attn2 of every block, place 'up', AttentionStore(img // 32, img // 8)) installed on its vendored blocks, the
    `vit-block{i}-self-map` / `-cross-map` ids through the real FeatureStore, a caption mask with padded tokens, and the
    aggregated `attn` feature assembled like diffusion_feature.py:488-500."""
    import torch.nn.functional as F
    cfg = TINY_DIT
    ref, omodel, sd = _build_ref_dit(cfg)
    rfe = ref_shim.load_reference_feature_extractor()
    rat = ref_shim.load_reference_attention_store()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.transformer = ref
    ids = _dit_feature_ids(cfg, with_maps=True)
    map_ids = [i for i in ids if i.endswith("-map")]
    L = cfg["sample_size"]
    img = 8 * L
    store = rfe.prepare_feature_extractor("pixart-sigma", pipe, {i: True for i in ids}, 1, True)
    astore = rat.register_attention_store("pixart-sigma", pipe, img, True)
    categories = ["up_cross", "up_self"]
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(1, 4, L, L, generator=g)
    _, ctx, mask, _, _ = make_dit_inputs(1, img, cfg["caption_dim"])
    mask = mask.clone()
    mask[:, -5:] = 0                                   # padded caption tokens: their probabilities must come out 0
    with torch.no_grad():
        out = ref(x, 50.0, ctx, mask)
    feats = dict(store.stored_feats)
    assert list(feats.keys()) == ids, list(feats.keys())[:10]
    all_attns = []
    for category, maps in astore.aggregate_attention(categories).items():
        for size, attn in maps.items():
            all_attns.append(F.interpolate(attn, size=(img // 8, img // 8)))
    attn_feat = torch.cat(all_attns, dim=-3)
    ostore = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers_dit(omodel, ostore)
    oast = O.register_attention_store_dit(omodel, img)
    with torch.no_grad():
        oout = omodel(x, 50.0, ctx, mask)
    assert list(ostore.feats.keys()) == ids
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in ids)
    oattn = O.aggregated_attention_feature(oast, categories, img)
    print("%s: %d ids (%d maps) + aggregated attn %s from the reference's AttnStoreProcessor on its vendored PixArt "
          "blocks; oracle max |diff| %.2e (attn %.2e, out %.2e)"
          % (name, len(ids), len(map_ids), tuple(attn_feat.shape), worst, (attn_feat - oattn).abs().max().item(),
             (out - oout).abs().max().item()))
    assert worst < 1e-3 and (attn_feat - oattn).abs().max().item() < 1e-5
    # drop the processors again (the oracle model object is shared with nothing else, but be explicit)
    torch.save({"ids": ids, "map_ids": map_ids, "categories": categories, "img": img, "x": x, "ctx": ctx, "mask": mask,
                "timestep": 50.0, "attn": attn_feat, "noise_pred": out,
                "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference AttnStoreProcessor / AttentionStore "
                             "on the vendored BasicTransformerBlock, norm_type ada_norm_single)"},
               os.path.join(OUT, name))


def golden_dit(name="dit_tiny_pixart.pt"):
    """PixArt path: the reference's vendored BasicTransformerBlock (attention.py:469-592, norm_type
    'ada_norm_single', attention_bias, gelu-approximate FeedForward) + vendored Attention / AttnProcessor2_0 +
    the reference's real prepare_feature_extractor (PixArt branch, feature_extractor.py:259-286) / FeatureStore.
    The outer model (PatchEmbed, AdaLayerNormSingle, caption projection, output head) is un-vendored in the
    reference and comes from the oracle's restatement -> those pieces stay PARITY UNPINNED."""
    cfg = TINY_DIT
    ref, omodel, sd = _build_ref_dit(cfg)
    rfe = ref_shim.load_reference_feature_extractor()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.transformer = ref
    ids = _dit_feature_ids(cfg)
    store = rfe.prepare_feature_extractor("pixart-sigma", pipe, {i: True for i in ids}, 1, True)
    g = torch.Generator().manual_seed(4321)
    L = cfg["sample_size"]
    x = torch.randn(1, 4, L, L, generator=g)
    _, ctx, mask, _, _ = make_dit_inputs(1, 8 * L, cfg["caption_dim"])
    ctx_b, mask_b = ctx, mask
    with torch.no_grad():
        out = ref(x, 50.0, ctx_b, mask_b)
    feats = store.stored_feats
    assert list(feats.keys()) == ids, list(feats.keys())[:8]
    ostore = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers_dit(omodel, ostore)
    with torch.no_grad():
        oout = omodel(x, 50.0, ctx_b, mask_b)
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in ids)
    print("%s: %d maps from the reference's vendored blocks; oracle max |diff| %.2e (out %.2e)"
          % (name, len(ids), worst, (out - oout).abs().max().item()))
    assert worst < 1e-3
    torch.save({"ids": ids, "x": x, "ctx": ctx, "mask": mask, "timestep": 50.0, "noise_pred": out,
                "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference vendored transformer blocks)"},
               os.path.join(OUT, name))


def golden_flux(name="flux_tiny.pt"):
    cfg = TINY_FLUX
    sd = models.synthetic_state_dict("flux", "cpu", None, TINY_VAE_FLUX, None, cfg)
    ref = ref_shim.build_reference_flux(cfg)
    ref.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)
    ref.eval()
    rfe = ref_shim.load_reference_feature_extractor()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.transformer = ref
    ids = _flux_feature_ids(cfg)
    store = rfe.prepare_feature_extractor("flux", pipe, {i: True for i in ids}, 1, True)
    img = 128
    L = img // 8
    g = torch.Generator().manual_seed(4321)
    lat = torch.randn(1, cfg["in_ch"] // 4, L, L, generator=g)     # x_t (already noised latents)
    x = O.flux_pack_latents(lat)
    _, ctx, pooled, _, _ = make_flux_inputs(1, img, cfg)
    sigma = O.resolve_flux_sigma(50, img)
    img_ids, txt_ids = O.flux_latent_image_ids(L // 2, L // 2), torch.zeros(cfg["ctx_len"], 3)
    with torch.no_grad():
        out = ref(hidden_states=x, encoder_hidden_states=ctx, pooled_projections=pooled,
                  timestep=torch.tensor([sigma]), img_ids=img_ids, txt_ids=txt_ids, guidance=torch.tensor([1.0]),
                  return_dict=False)[0]
    feats = store.stored_feats
    assert list(feats.keys()) == ids, list(feats.keys())[:8]
    omodel, _ = build_oracle_flux(cfg, TINY_VAE_FLUX, sd)
    ostore = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers_flux(omodel, ostore)
    with torch.no_grad():
        oout = omodel(x, ctx, pooled, sigma, img_ids, txt_ids, 1.0)
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in ids)
    print("%s: %d maps from the reference's vendored Flux transformer; oracle max |diff| %.2e (out %.2e)"
          % (name, len(ids), worst, (out - oout).abs().max().item()))
    assert worst < 1e-3
    torch.save({"ids": ids, "latents": lat, "ctx": ctx, "pooled": pooled, "sigma": sigma, "guidance": 1.0,
                "noise_pred": out, "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference vendored transformer_flux.py)"},
               os.path.join(OUT, name))


def golden_flux_maps(name="flux_tiny_maps.pt"):
    """Attention-probability maps of the Flux family: the reference's REAL FluxAttnStoreProcessor /
    AttentionStore / register_attention_store (feature/components/attention.py:402-527, 567-603) installed on its whole
    vendored FluxTransformer2DModel: per block `cross-map` (image queries x text keys) and `self-map` (image x image)
    through the real FeatureStore + the aggregated `attn` feature (diffusion_feature.py:488-500)."""
    import torch.nn.functional as F
    cfg = TINY_FLUX
    sd = models.synthetic_state_dict("flux", "cpu", None, TINY_VAE_FLUX, None, cfg)
    ref = ref_shim.build_reference_flux(cfg)
    ref.load_state_dict({k[len("transformer."):]: v for k, v in sd.items() if k.startswith("transformer.")}, strict=True)
    ref.eval()
    rfe = ref_shim.load_reference_feature_extractor()
    rat = ref_shim.load_reference_attention_store()

    class Pipe:
        pass
    pipe = Pipe()
    pipe.transformer = ref
    ids = _flux_feature_ids(cfg, with_maps=True)
    map_ids = [i for i in ids if i.endswith("-map")]
    img = 128
    store = rfe.prepare_feature_extractor("flux", pipe, {i: True for i in ids}, 1, True)
    astore = rat.register_attention_store("flux", pipe, img, True)
    categories = ["up_cross", "up_self"]
    L = img // 8
    g = torch.Generator().manual_seed(4321)
    lat = torch.randn(1, cfg["in_ch"] // 4, L, L, generator=g)
    x = O.flux_pack_latents(lat)
    _, ctx, pooled, _, _ = make_flux_inputs(1, img, cfg)
    sigma = O.resolve_flux_sigma(50, img)
    img_ids, txt_ids = O.flux_latent_image_ids(L // 2, L // 2), torch.zeros(cfg["ctx_len"], 3)
    with torch.no_grad():
        out = ref(hidden_states=x, encoder_hidden_states=ctx, pooled_projections=pooled,
                  timestep=torch.tensor([sigma]), img_ids=img_ids, txt_ids=txt_ids, guidance=torch.tensor([1.0]),
                  return_dict=False)[0]
    feats = dict(store.stored_feats)
    assert list(feats.keys()) == ids, list(feats.keys())[:12]
    all_attns = []
    for category, maps in astore.aggregate_attention(categories).items():
        for size, attn in maps.items():
            all_attns.append(F.interpolate(attn, size=(img // 8, img // 8)))
    attn_feat = torch.cat(all_attns, dim=-3)
    omodel, _ = build_oracle_flux(cfg, TINY_VAE_FLUX, sd)
    ostore = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers_flux(omodel, ostore)
    oast = O.register_attention_store_flux(omodel, img)
    with torch.no_grad():
        oout = omodel(x, ctx, pooled, sigma, img_ids, txt_ids, 1.0)
    assert list(ostore.feats.keys()) == ids
    worst = max((feats[k] - ostore.feats[k]).abs().max().item() for k in ids)
    oattn = O.aggregated_attention_feature(oast, categories, img)
    print("%s: %d ids (%d maps) + aggregated attn %s from the reference's FluxAttnStoreProcessor; oracle max |diff| %.2e "
          "(attn %.2e, out %.2e)" % (name, len(ids), len(map_ids), tuple(attn_feat.shape), worst,
                                     (attn_feat - oattn).abs().max().item(), (out - oout).abs().max().item()))
    assert worst < 1e-3 and (attn_feat - oattn).abs().max().item() < 1e-5
    torch.save({"ids": ids, "map_ids": map_ids, "categories": categories, "img": img, "latents": lat, "ctx": ctx,
                "pooled": pooled, "sigma": sigma, "guidance": 1.0, "attn": attn_feat, "noise_pred": out,
                "feats": {k: v.to(torch.float16) for k, v in feats.items()},
                "generator": "tools/make_golden.py via tools/ref_shim.py (reference FluxAttnStoreProcessor on the vendored "
                             "transformer_flux.py)"},
               os.path.join(OUT, name))


def golden_store_resize(name="feature_store_resize.pt"):
    """feature_resize: the reference's REAL FeatureStore.store (feature/components/feature_extractor.py:31-76,
    adaptive_avg_pool2d at :51-53) on seeded conv-style (B,C,h,w) and ViT-style (B,N,C) activations, ratios 2 and 3
    (3 does not divide 16: uneven adaptive windows)."""
    rfe = ref_shim.load_reference_feature_extractor()
    g = torch.Generator().manual_seed(77)
    conv = torch.randn(2, 64, 16, 16, generator=g)
    vit = torch.randn(2, 256, 128, generator=g)
    out = {"conv": conv, "vit": vit}
    for r in (2, 3):
        st = rfe.FeatureStore({"a": True, "b": True}, r, True)
        st.store(conv, "a")
        st.store(vit, "b")
        out["r%d" % r] = {k: v.clone() for k, v in st.stored_feats.items()}
        ost = O.FeatureStore({"a": True, "b": True}, r)
        ost.store(conv, "a")
        ost.store(vit, "b")
        for k in ("a", "b"):
            assert torch.allclose(ost.feats[k], out["r%d" % r][k], atol=1e-6), (r, k)
    out["generator"] = "tools/make_golden.py (reference FeatureStore.store)"
    torch.save(out, os.path.join(OUT, name))
    print("%s: reference FeatureStore.store with resize_ratio 2 / 3; oracle identical" % name)


def golden_cli(name="cli_layouts.json"):
    """On-disk format: the reference's OWN extract_feature.py main() (argument parsing, naming, directory layout,
    aggregation with F.interpolate + cat, np.save) executed on a stand-in FeatureExtractor; the file trees (paths,
    shapes, dtypes, content hashes) of three flag combinations are the fixture."""
    import importlib.util
    import json
    import tempfile
    import types
    fake_mod = types.ModuleType("diffusion_feature")
    fake_mod.FeatureExtractor = FakeExtractor
    sys.modules["diffusion_feature"] = fake_mod
    spec = importlib.util.spec_from_file_location("ref_extract_feature", os.path.join(ref_shim.REF, "extract_feature.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        cli_fixture_inputs(root)
        for mode, flags in CLI_MODES.items():
            od = os.path.join(root, "out_" + mode)
            argv = ["extract_feature.py", "--layer", "x.json", "--t", "50", "-b", "4", "--input_dir",
                    os.path.join(root, "imgs", "*", "*.png"), "--prompt_file", os.path.join(root, "prompt.txt"),
                    "--output_dir", od] + flags
            old = sys.argv
            sys.argv = argv
            try:
                ref.main()
            finally:
                sys.argv = old
            out[mode] = tree_digest(od)
    del sys.modules["diffusion_feature"]
    json.dump(out, open(os.path.join(OUT, name), "w"), indent=0, sort_keys=True)
    print("%s: %s files" % (name, {k: len(v) for k, v in out.items()}))


def golden_correspondence():
    cu = ref_shim.load_reference_correspondence_utils()
    g = torch.Generator().manual_seed(99)
    f1 = torch.randn(1, 64, 16, 16, generator=g)
    f2 = torch.randn(1, 64, 16, 16, generator=g)
    rng = np.random.RandomState(5)
    pts = rng.uniform(-2, 66, size=(40, 2))           # some outside [0, 63] to exercise the clip
    load_size = (64, 64)
    p1, p2 = cu.find_nn_source_correspondences(f1, f2, pts, None, load_size)
    idx = cu.points_to_idxs(pts, load_size)
    o_p2, _ = O.find_nn_source_correspondences(f1, f2, pts, load_size)
    assert torch.equal(p2, o_p2), "oracle correspondence differs from the reference's"
    assert np.array_equal(idx, O.points_to_idxs(pts, load_size))
    print("correspondence.pt: %d points, oracle == reference" % len(pts))
    torch.save({"f1": f1, "f2": f2, "points": torch.from_numpy(pts), "load_size": load_size, "points2": p2,
                "idx": torch.from_numpy(idx)}, os.path.join(OUT, "correspondence.pt"))


def _seed_head(m, g):
    for n, p in m.named_parameters():
        with torch.no_grad():
            if n.endswith(".0.weight"):     # fp16-representable, stored as fp16 (fixture size)
                p.copy_((torch.randn(p.shape, generator=g) * (0.5 / (9 * p.shape[1]) ** 0.5)).half().float())
            elif n.endswith(".1.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    for n, b in m.named_buffers():
        if n.endswith("running_mean"):
            b.copy_(0.1 * torch.randn(b.shape, generator=g))
        elif n.endswith("running_var"):
            b.copy_(0.5 + torch.rand(b.shape, generator=g))


def _golden_segmentor_multi(seg, g):
    """Several-extractors branch of the REAL DiffusionSegmentor.extract_feat (diffusion_segmentor.py:248-297): two stand-in
    extractors, MultiRes(dim, 4) per map, MultiRes(sum, 2) per model and level, ResBlock 'amalgemated' per level."""
    layers = [[[("up-level0-upsampler-out", 64)], [("up-level1-upsampler-out", 64)]],
              [[("mid-vit-out", 64)], []]]
    c_per_level = [128, 64]
    sizes = [8, 16]
    B = 1
    feats = [{}, {}]
    for i, ls in enumerate(layers):
        for level, res in enumerate(ls):
            for lname, c in res:
                feats[i][lname] = torch.randn(B, c, sizes[level], sizes[level], generator=g).to(torch.float16)

    class _FE:
        def __init__(self, i):
            self.i = i

        def extract(self, **kw):
            assert kw["image_type"] == "tensors"
            return feats[self.i]

    m = object.__new__(seg.DiffusionSegmentor)
    torch.nn.Module.__init__(m)
    m.multiple_diffusion = True
    m.feature_extractors = [{"model": _FE(i), "prompt_embeds": None, "t": 50, "layers": ls} for i, ls in enumerate(layers)]
    for i, ls in enumerate(layers):
        for rank, res in enumerate(ls):
            for lname, c in res:
                setattr(m, m.layer_conv_name(lname, i), seg.MultiRes(c, 4))
            if res:
                setattr(m, m.layer_conv_name("sum%d" % rank, i), seg.MultiRes(sum(c for _, c in res), 2))
    for i, dim in enumerate(c_per_level):
        setattr(m, m.layer_conv_name("amalgemated", i), seg.ResBlock(dim))
    _seed_head(m, g)
    m.eval()
    with torch.no_grad():
        outs = m.extract_feat(torch.zeros(B, 3, 8, 8), is_test=True)
    sd = {k: v.clone() for k, v in m.state_dict().items() if "num_batches_tracked" not in k and ".res." not in k
          or ".res.0." in k}
    sd = {k: v for k, v in sd.items() if "num_batches_tracked" not in k}
    o_outs = O.seg_extract_feat_multi(feats, layers, c_per_level, sd)
    for a, b in zip(outs, o_outs):
        assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), "oracle several-extractors segmentor head differs from the reference's"
    sd = {k: (v.half() if k.endswith(".0.weight") else v) for k, v in sd.items()}
    print("segmentor_head.pt[multi]: %s, oracle == reference" % [tuple(o.shape) for o in outs])
    return {"feature_layers": layers, "c_per_level": c_per_level, "features": feats, "state_dict": sd,
            "outs": [o.half() for o in outs]}


def golden_segmentor(name="segmentor_head.pt"):
    """Real ResBlock / DiffusionSegmentor.extract_feat of segmentation/models/diffusion_segmentor.py (single-extractor
    branch, eval mode) on seeded fp16 maps. The reference initialises every ResBlock parameter to zero (an identity at
    the start of training), so seeded weights and BatchNorm running statistics stand in for a trained head."""
    seg = ref_shim.load_reference_segmentor()
    feature_layers = [[("up-level0-upsampler-out", 64)],
                      [("up-level1-upsampler-out", 64), ("up-level2-repeat2-res-out", 64)]]
    sizes = [16, 32]
    g = torch.Generator().manual_seed(2718)
    B = 2
    feats = {}
    for level, res in enumerate(feature_layers):
        for lname, c in res:
            feats[lname] = torch.randn(B, c, sizes[level], sizes[level], generator=g).to(torch.float16)

    class _FE:
        def extract(self, **kw):
            assert kw["image_type"] == "tensors" and kw["t"] == 50
            return feats

    m = object.__new__(seg.DiffusionSegmentor)
    torch.nn.Module.__init__(m)
    m.multiple_diffusion = False
    m.t = 50
    m.feature_layers = feature_layers
    m.prompt_embeds = None
    m.feature_extractor = _FE()
    for rank, res in enumerate(feature_layers):
        for lname, c in res:
            setattr(m, m.layer_conv_name(lname), seg.ResBlock(c))
        setattr(m, m.layer_conv_name("sum%d" % rank), seg.ResBlock(sum(c for _, c in res)))
    for n, p in m.named_parameters():
        with torch.no_grad():
            if n.endswith(".0.weight"):     # fp16-representable, stored as fp16 (fixture size)
                p.copy_((torch.randn(p.shape, generator=g) * (0.7 / (9 * p.shape[1]) ** 0.5)).half().float())
            elif n.endswith(".1.weight"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    for n, b in m.named_buffers():
        if n.endswith("running_mean"):
            b.copy_(0.1 * torch.randn(b.shape, generator=g))
        elif n.endswith("running_var"):
            b.copy_(0.5 + torch.rand(b.shape, generator=g))
    m.eval()
    with torch.no_grad():
        outs = m.extract_feat(torch.zeros(B, 3, 8, 8), is_test=True)
    sd = {k: v.clone() for k, v in m.state_dict().items() if "num_batches_tracked" not in k}
    o_outs = O.seg_extract_feat(feats, feature_layers, sd)
    for a, b in zip(outs, o_outs):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5), "oracle segmentor head differs from the reference's"
    # MultiRes (diffusion_segmentor.py:46-53): one shared ResBlock applied n times
    mr = seg.MultiRes(64, 3)
    mr.load_state_dict({"res.%d.%s" % (i, k[len("up_level1_upsampler_out."):]): v for i in range(3)
                        for k, v in sd.items() if k.startswith("up_level1_upsampler_out.")}, strict=False)
    mr.eval()
    with torch.no_grad():
        mr_out = mr(feats["up-level1-upsampler-out"].float())
    sd = {k: (v.half() if k.endswith(".0.weight") else v) for k, v in sd.items()}
    multi = _golden_segmentor_multi(seg, g)
    torch.save({"multi": multi, "feature_layers": feature_layers, "features": feats, "state_dict": sd,
                "outs": [o.half() for o in outs], "multires_n": 3, "multires_out": mr_out.half()}, os.path.join(OUT, name))
    print("%s: %d levels %s, oracle == reference" % (name, len(outs), [tuple(o.shape) for o in outs]))


def golden_extract():
    sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    unet, vae = build_oracle(TINY_XL, TINY_VAE, sd)
    ids = _unet_feature_ids(TINY_XL)
    store = O.FeatureStore({i: True for i in ids})
    O.attach_gatherers(unet, store)
    image, ctx, pooled, ev, eq = make_inputs(1, 64, TINY_XL["ctx_dim"], 64)
    feats, latents, npred = O.extract("xl", unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=64)
    torch.save({"ids": ids, "latents": latents, "noise_pred": npred,
                "stats": {k: (float(v.mean()), float(v.std()), float(v.abs().max())) for k, v in feats.items()},
                "feats_subset": {k: feats[k].to(torch.float16) for k in ids[::12]},
                "generator": "tools/make_golden.py (oracle self-consistency; VAE / schedulers un-vendored)"},
               os.path.join(OUT, "extract_tiny_xl.pt"))
    print("extract_tiny_xl.pt: %d maps (oracle)" % len(ids))


def golden_ids():
    """Id sets and channel sums straight from the reference's own JSON artefacts."""
    import json
    ref = ref_shim.REF
    out = {}
    for key, f in (("xl", "config_xl_full.json"), ("1-5", "config_15_full.json")):
        cfg = json.load(open(os.path.join(ref, "feature", "configs", f)))
        out["ids_" + key] = [k for k in cfg if "map" not in k]
        out["map_ids_" + key] = [k for k in cfg if "map" in k]
        out["all_ids_" + key] = list(cfg)                  # file order = execution order, maps interleaved
    for f in ("config_xl_practical.json", "config_xl_legacy.json", "config_15_practical.json",
              "config_15_legacy.json"):
        out[f] = json.load(open(os.path.join(ref, "feature", "configs", f)))
    for f in ("config_sdxl.json", "config_sd15.json", "config_legacy_sdxl.json", "config_legacy_sd15.json"):
        c = json.load(open(os.path.join(ref, "correspondence", "correspondence", f)))
        out["corr_" + f] = {"layer": c["layer"], "version": c["version"], "feature_len": c["feature_len"],
                            "img_size": c["img_size"], "t": c["t"]}
    json.dump(out, open(os.path.join(OUT, "reference_ids.json"), "w"), indent=0)
    print("reference_ids.json: %d xl ids, %d 1-5 ids" % (len(out["ids_xl"]), len(out["ids_1-5"])))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_ids()
    golden_unet("xl", TINY_XL, "unet_tiny_xl.pt")
    golden_unet("2-1", TINY_21, "unet_tiny_21.pt")
    golden_unet("1-5", TINY_15, "unet_tiny_15.pt")       # conv proj_in / proj_out (use_linear_projection False), 8 heads
    golden_unet_maps()
    golden_dit_maps()
    golden_flux_maps()
    golden_unet_control()
    golden_dit()
    golden_flux()
    golden_store_resize()
    golden_cli()
    golden_correspondence()
    golden_segmentor()
    golden_extract()
