"""CPU-only tests (`-m "not gpu"`): the oracle against the golden fixtures produced from the reference's own code,
the host logic (id grammar, parameter naming, schedulers, error behaviour, sharding) and the C-ABI surface
(library loads, exports every symbol include/gdf.h declares; no compute calls)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import (O, ROOT, TINY_15, TINY_21, TINY_DIT, TINY_FLUX, TINY_VAE, TINY_VAE_FLUX, TINY_XL, build_oracle, build_oracle_dit,
                    build_oracle_flux, make_inputs)

GOLD = os.path.join(ROOT, "tests", "golden")


def _models():
    from generic_diffusion_feature_b200.components import models
    return models


# ------------------------------------------------------------------------------------------ oracle vs golden
@pytest.mark.parametrize("fixture,version,cfg", [("unet_tiny_xl.pt", "xl", TINY_XL), ("unet_tiny_21.pt", "2-1", TINY_21),
                                                 ("unet_tiny_15.pt", "1-5", TINY_15)])
def test_oracle_matches_reference_vendored_unet(fixture, version, cfg):
    """tests/golden/unet_tiny_*.pt were produced by the reference's vendored UNet2DConditionModel + its own
    FeatureStore (tools/make_golden.py); the oracle must reproduce every map (fixtures are fp16-rounded)."""
    gold = torch.load(os.path.join(GOLD, fixture), weights_only=False)
    sd = _models().synthetic_state_dict(version, "cpu", cfg, TINY_VAE)
    unet, _ = build_oracle(cfg, TINY_VAE, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers(unet, store)
    kw = {}
    if gold["pooled"] is not None:
        kw = dict(text_embeds=gold["pooled"], time_ids=O.add_time_ids(8 * gold["x"].shape[-1]))
    with torch.no_grad():
        out = unet(gold["x"], gold["timestep"], gold["ctx"], **kw)
    assert list(store.feats.keys()) == gold["ids"]
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        got = store.feats[k]
        assert got.shape == ref.shape, k
        tol = 2e-3 * max(1.0, ref.abs().max().item())      # fp16 rounding of the stored fixture
        assert (got - ref).abs().max().item() <= tol, k


def test_oracle_controlnet_residual_inputs_match_reference():
    """tests/golden/unet_tiny_xl_control.pt (oracle side; the CUDA path does not take these inputs
    yet): the reference's vendored UNet forward with down_block_additional_residuals / mid_block_additional_residual
    (unet_2d_condition.py:1236-1275)."""
    gold = torch.load(os.path.join(GOLD, "unet_tiny_xl_control.pt"), weights_only=False)
    sd = _models().synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    unet, _ = build_oracle(TINY_XL, TINY_VAE, sd)
    tid = O.add_time_ids(8 * gold["x"].shape[-1])
    with torch.no_grad():
        out = unet(gold["x"], gold["timestep"], gold["ctx"], text_embeds=gold["pooled"], time_ids=tid,
                   down_residuals=gold["down"], mid_residual=gold["mid"])
        plain = unet(gold["x"], gold["timestep"], gold["ctx"], text_embeds=gold["pooled"], time_ids=tid)
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    assert (plain - gold["noise_pred"]).abs().max().item() > 1e-2          # the residuals matter


def test_oracle_attention_maps_match_reference_processor():
    """tests/golden/unet_tiny_xl_maps.pt (oracle side): the reference's real AttnStoreProcessor /
    AttentionStore / register_attention_store on its vendored UNet - every `...-self-map` / `...-cross-map`
    (B, heads, Nq, Nk), the id order with the maps interleaved, and the aggregated `attn` feature of
    diffusion_feature.py:488-500."""
    gold = torch.load(os.path.join(GOLD, "unet_tiny_xl_maps.pt"), weights_only=False)
    sd = _models().synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    unet, _ = build_oracle(TINY_XL, TINY_VAE, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers(unet, store)
    ast = O.register_attention_store(unet, gold["img"])
    with torch.no_grad():
        unet(gold["x"], gold["timestep"], gold["ctx"], text_embeds=gold["pooled"], time_ids=O.add_time_ids(gold["img"]))
    assert list(store.feats.keys()) == gold["ids"] and len(gold["map_ids"]) == 34
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        tol = 2e-3 * max(1.0, ref.abs().max().item())
        assert store.feats[k].shape == ref.shape and (store.feats[k] - ref).abs().max().item() <= tol, k
    m = store.feats["mid-vit-block0-cross-map"]
    assert m.dim() == 4 and m.shape[1] == TINY_XL["heads"][-1] and m.shape[-1] == 77
    assert torch.allclose(m.sum(-1), torch.ones_like(m.sum(-1)), atol=1e-5)          # rows are probabilities
    attn = O.aggregated_attention_feature(ast, gold["categories"], gold["img"])
    assert attn.shape == gold["attn"].shape and (attn - gold["attn"]).abs().max().item() < 1e-5


def test_oracle_pixart_attention_maps_match_reference_processor():
    """tests/golden/dit_tiny_pixart_maps.pt (transformer pipes): the reference's real
    AttnStoreProcessor / AttentionStore / register_attention_store (feature/components/attention.py:567-593: attn1 and
    attn2 of every block, place 'up', AttentionStore(img // 32, img // 8)) on its vendored ada_norm_single blocks with a
    caption mask that pads 5 tokens - every `vit-block{i}-self-map` / `-cross-map`, the id order, the aggregated `attn`
    feature; and the host-side mirror (attention_mean_ids / aggregate_attention) fed with the oracle's head means."""
    from generic_diffusion_feature_b200.components.feature_extractor import (ATTN_MEAN_PREFIX, _dit_feature_ids,
                                                                             aggregate_attention, attention_mean_ids)
    gold = torch.load(os.path.join(GOLD, "dit_tiny_pixart_maps.pt"), weights_only=False)
    assert gold["ids"] == _dit_feature_ids(TINY_DIT, with_maps=True)
    sd = _models().synthetic_state_dict("pixart-sigma", "cpu", None, TINY_VAE, TINY_DIT)
    model, _ = build_oracle_dit(TINY_DIT, TINY_VAE, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers_dit(model, store)
    ast = O.register_attention_store_dit(model, gold["img"])
    with torch.no_grad():
        out = model(gold["x"], gold["timestep"], gold["ctx"], gold["mask"])
    assert list(store.feats.keys()) == gold["ids"] and len(gold["map_ids"]) == 2 * TINY_DIT["layers"]
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        tol = 2e-3 * max(1.0, ref.abs().max().item())
        assert store.feats[k].shape == ref.shape and (store.feats[k] - ref).abs().max().item() <= tol, k
    m = store.feats["vit-block0-cross-map"]
    assert m.dim() == 4 and m.shape[1] == TINY_DIT["heads"]
    assert torch.allclose(m.sum(-1), torch.ones_like(m.sum(-1)), atol=1e-5)
    assert m[..., -5:].abs().max().item() == 0.0                      # padded caption tokens get probability 0
    attn = O.aggregated_attention_feature(ast, gold["categories"], gold["img"])
    assert attn.shape == gold["attn"].shape and (attn - gold["attn"]).abs().max().item() < 1e-5
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    # host mirror: plan-order head means -> the same `attn`
    mean_ids = attention_mean_ids(None, gold["categories"], dit_cfg=TINY_DIT)
    assert mean_ids == [ATTN_MEAN_PREFIX + "vit-block%d-%s" % (i, k) for i in range(TINY_DIT["layers"])
                        for k in ("self", "cross")]
    means = [(i[len(ATTN_MEAN_PREFIX):].rsplit("-", 1)[0], i.rsplit("-", 1)[1],
              store.feats[i[len(ATTN_MEAN_PREFIX):] + "-map"].mean(1)) for i in mean_ids]
    host = aggregate_attention(means, gold["categories"], gold["img"], transformer=True)
    assert host.shape == gold["attn"].shape and (host.float() - gold["attn"]).abs().max().item() < 2e-3


def test_feature_plan_views_layouts():
    """FeaturePlan.views / attention_means on a hand-made slot table (pure tensor views, no GPU): token-major maps come
    back as (B, C, h, w) with channel stride 1, `...-map` slots as the contiguous (B, heads, Nq, Nk) tensor the
    reference stores, internal `#attnmean:` slots only through attention_means, everything in execution order."""
    from generic_diffusion_feature_b200.components.feature_extractor import ATTN_MEAN_PREFIX, FeaturePlan
    plan = FeaturePlan.__new__(FeaturePlan)
    plan.batch = 2
    B = 2
    # (id, offset, C, H, W, order): a 3-channel 2x2 map, a (heads=2, Nq=4, Nk=3) probability map, its head mean
    n0, n1, n2 = B * 2 * 2 * 3, B * 2 * 4 * 3, B * 4 * 3
    plan.slots = [("mid-vit-block0-self-map", 256, 2, 4, 3, 1), ("mid-vit-out", 0, 3, 2, 2, 3),
                  (ATTN_MEAN_PREFIX + "mid-vit-block0-self", 512, 1, 4, 3, 2), ("mid-vit-block0-cross-k", -1, 0, 0, 0, -1)]
    arena = torch.zeros(1024, dtype=torch.uint8)
    arena[0:2 * n0].view(torch.float16)[:] = torch.arange(n0, dtype=torch.float16)
    arena[256:256 + 2 * n1].view(torch.float16)[:] = torch.arange(n1, dtype=torch.float16) + 100
    arena[512:512 + 2 * n2].view(torch.float16)[:] = torch.arange(n2, dtype=torch.float16) + 500
    views = plan.views(arena)
    assert list(views.keys()) == ["mid-vit-block0-self-map", "mid-vit-out"]            # execution order, internal ids hidden
    v = views["mid-vit-out"]
    assert v.shape == (2, 3, 2, 2) and v.stride(1) == 1 and v[1, 2, 1, 0].item() == float((1 * 4 + 2) * 3 + 2)
    m = views["mid-vit-block0-self-map"]
    assert m.shape == (2, 2, 4, 3) and m.is_contiguous() and m[1, 1, 3, 2].item() == 100.0 + n1 - 1
    means = plan.attention_means(arena)
    assert [(b, k) for b, k, _ in means] == [("mid-vit-block0", "self")]
    assert means[0][2].shape == (2, 4, 3) and means[0][2][1, 3, 2].item() == 500.0 + n2 - 1


def test_host_attention_aggregation_matches_reference():
    """Host side of the aggregated `attn` feature (attention_mean_ids + aggregate_attention, the mirror of
    register_attention_store / AttentionStore.aggregate_attention / diffusion_feature.py:488-500) fed with the oracle's
    head-mean maps in plan order reproduces the reference fixture."""
    from generic_diffusion_feature_b200.components.feature_extractor import (ATTN_MEAN_PREFIX, aggregate_attention,
                                                                             attention_mean_ids)
    gold = torch.load(os.path.join(GOLD, "unet_tiny_xl_maps.pt"), weights_only=False)
    sd = _models().synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    unet, _ = build_oracle(TINY_XL, TINY_VAE, sd)
    ast = O.register_attention_store(unet, gold["img"])
    with torch.no_grad():
        unet(gold["x"], gold["timestep"], gold["ctx"], text_embeds=gold["pooled"], time_ids=O.add_time_ids(gold["img"]))
    ids = attention_mean_ids(TINY_XL, gold["categories"])
    assert ids and all(i.startswith(ATTN_MEAN_PREFIX) for i in ids)
    pools = {k: list(v) for k, v in ast.step_store.items()}
    means = []
    for i in ids:                                   # plan (= execution) order; the store keeps per-category lists
        block, kind = i[len(ATTN_MEAN_PREFIX):].rsplit("-", 1)
        means.append((block, kind, pools["%s_%s" % (block.split("-")[0], kind)].pop(0).to(torch.float16)))
    attn = aggregate_attention(means, gold["categories"], gold["img"])
    assert attn.shape == gold["attn"].shape and attn.dtype == torch.float16
    assert (attn.float() - gold["attn"]).abs().max().item() < 2e-3
    # every category the CLI accepts maps to at least one attention module of the SDXL UNet
    full = _models().UNET_CONFIGS["xl"]
    for cat in ("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self"):
        assert attention_mean_ids(full, [cat])
    assert len(attention_mean_ids(full, ["down_cross", "mid_cross", "up_cross"])) == 70


def test_oracle_matches_reference_vendored_dit_blocks():
    """tests/golden/dit_tiny_pixart.pt: the reference's vendored BasicTransformerBlock (ada_norm_single) stack +
    its real prepare_feature_extractor PixArt branch (tools/make_golden.py); the oracle reproduces every map."""
    from generic_diffusion_feature_b200.components.feature_extractor import _dit_feature_ids
    gold = torch.load(os.path.join(GOLD, "dit_tiny_pixart.pt"), weights_only=False)
    assert gold["ids"] == _dit_feature_ids(TINY_DIT)
    sd = _models().synthetic_state_dict("pixart-sigma", "cpu", None, TINY_VAE, TINY_DIT)
    model, _ = build_oracle_dit(TINY_DIT, TINY_VAE, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers_dit(model, store)
    with torch.no_grad():
        out = model(gold["x"], gold["timestep"], gold["ctx"], gold["mask"])
    assert list(store.feats.keys()) == gold["ids"]
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        tol = 2e-3 * max(1.0, ref.abs().max().item())
        assert store.feats[k].shape == ref.shape and (store.feats[k] - ref).abs().max().item() <= tol, k


def test_oracle_matches_reference_vendored_flux():
    """tests/golden/flux_tiny.pt: the reference's vendored FluxTransformer2DModel (transformer_flux.py ctor + forward,
    double / single blocks, FluxAttnProcessor2_0) + its real prepare_feature_extractor Flux branch
    (tools/make_golden.py); the oracle reproduces every map and the id order."""
    from generic_diffusion_feature_b200.components.feature_extractor import _flux_feature_ids
    gold = torch.load(os.path.join(GOLD, "flux_tiny.pt"), weights_only=False)
    assert gold["ids"] == _flux_feature_ids(TINY_FLUX)
    sd = _models().synthetic_state_dict("flux", "cpu", None, TINY_VAE_FLUX, None, TINY_FLUX)
    model, _ = build_oracle_flux(TINY_FLUX, TINY_VAE_FLUX, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers_flux(model, store)
    L = gold["latents"].shape[-1]
    with torch.no_grad():
        out = model(O.flux_pack_latents(gold["latents"]), gold["ctx"], gold["pooled"], gold["sigma"],
                    O.flux_latent_image_ids(L // 2, L // 2), torch.zeros(gold["ctx"].shape[1], 3), gold["guidance"])
    assert list(store.feats.keys()) == gold["ids"]
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        tol = 2e-3 * max(1.0, ref.abs().max().item())
        assert store.feats[k].shape == ref.shape and (store.feats[k] - ref).abs().max().item() <= tol, k
    # quirk kept from the reference (transformer_flux.py:200-211): `out` of a double block stores norm_hidden_states
    assert torch.equal(store.feats["vit-block0-out"], store.feats["vit-block0-norm-out"])


def test_oracle_flux_attention_maps_match_reference_processor():
    """tests/golden/flux_tiny_maps.pt (MMDiT): the reference's real FluxAttnStoreProcessor /
    AttentionStore / register_attention_store (feature/components/attention.py:402-527, 567-603) on its whole vendored
    FluxTransformer2DModel - per block `cross-map` (image queries x text keys) and `self-map` (image x image), gathered
    in that order right after q / k / v, and the aggregated `attn` feature; plus the host-side mirror."""
    from generic_diffusion_feature_b200.components.feature_extractor import (ATTN_MEAN_PREFIX, _flux_feature_ids,
                                                                             aggregate_attention, attention_mean_ids)
    gold = torch.load(os.path.join(GOLD, "flux_tiny_maps.pt"), weights_only=False)
    assert gold["ids"] == _flux_feature_ids(TINY_FLUX, with_maps=True)
    sd = _models().synthetic_state_dict("flux", "cpu", None, TINY_VAE_FLUX, None, TINY_FLUX)
    model, _ = build_oracle_flux(TINY_FLUX, TINY_VAE_FLUX, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers_flux(model, store)
    ast = O.register_attention_store_flux(model, gold["img"])
    L = gold["latents"].shape[-1]
    with torch.no_grad():
        out = model(O.flux_pack_latents(gold["latents"]), gold["ctx"], gold["pooled"], gold["sigma"],
                    O.flux_latent_image_ids(L // 2, L // 2), torch.zeros(gold["ctx"].shape[1], 3), gold["guidance"])
    assert list(store.feats.keys()) == gold["ids"]
    assert (out - gold["noise_pred"]).abs().max().item() < 1e-4
    for k in gold["ids"]:
        ref = gold["feats"][k].float()
        tol = 2e-3 * max(1.0, ref.abs().max().item())
        assert store.feats[k].shape == ref.shape and (store.feats[k] - ref).abs().max().item() <= tol, k
    n_img, n_txt = (L // 2) ** 2, TINY_FLUX["ctx_len"]
    c, m = store.feats["vit-block0-cross-map"], store.feats["vit-block0-self-map"]
    assert c.shape == (1, TINY_FLUX["heads"], n_img, n_txt) and m.shape == (1, TINY_FLUX["heads"], n_img, n_img)
    assert torch.allclose(c.sum(-1) + m.sum(-1), torch.ones_like(c.sum(-1)), atol=1e-5)   # one softmax over both parts
    attn = O.aggregated_attention_feature(ast, gold["categories"], gold["img"])
    assert attn.shape == gold["attn"].shape and (attn - gold["attn"]).abs().max().item() < 1e-5
    mean_ids = attention_mean_ids(None, gold["categories"], flux_cfg=TINY_FLUX)
    nb = TINY_FLUX["layers"] + TINY_FLUX["single_layers"]
    assert mean_ids == [ATTN_MEAN_PREFIX + "vit-block%d-%s" % (i, k) for i in range(nb) for k in ("cross", "self")]
    means = [(i[len(ATTN_MEAN_PREFIX):].rsplit("-", 1)[0], i.rsplit("-", 1)[1],
              store.feats[i[len(ATTN_MEAN_PREFIX):] + "-map"].mean(1)) for i in mean_ids]
    host = aggregate_attention(means, gold["categories"], gold["img"], transformer=True)
    assert host.shape == gold["attn"].shape and (host.float() - gold["attn"]).abs().max().item() < 2e-3


def test_flux_host_logic_matches_oracle():
    """Parameter naming, rotary tables, the resolved flow-match sigma and the id grammar of the host side agree with
    the oracle's restatement (FluxTransformer2DModel state_dict, FluxPosEmbed, pipeline_flux_img2img.py:745-766)."""
    m = _models()
    from generic_diffusion_feature_b200 import schedulers
    from generic_diffusion_feature_b200.components.feature_extractor import _flux_feature_ids
    for cfg in (TINY_FLUX, dict(m.FLUX_CONFIGS["flux"], layers=1, single_layers=1)):
        want = {k: tuple(v.shape) for k, v in O.FluxTransformer2DModel(cfg).state_dict().items()} \
            if cfg is TINY_FLUX else None
        if want is not None:
            assert dict(m.flux_param_specs(cfg)) == want
    full = m.FLUX_CONFIGS["flux"]
    n = sum(int(np.prod(s)) for _, s in m.flux_param_specs(full))
    assert 11.8e9 < n < 12.0e9                       # FLUX.1-dev: 11.9 B transformer parameters
    ids = _flux_feature_ids(full)
    assert len(ids) == 19 * 7 + 38 * 5 and ids[0] == "vit-block0-q" and ids[-1] == "vit-block56-out"
    cos, sin = m.flux_rope_tables(TINY_FLUX, TINY_FLUX["ctx_len"], 8)
    ids3 = torch.cat([torch.zeros(TINY_FLUX["ctx_len"], 3), O.flux_latent_image_ids(8, 8)], dim=0)
    ocos, osin = O.flux_rope(ids3, TINY_FLUX["axes_dims_rope"])
    assert torch.equal(cos, ocos) and torch.equal(sin, osin)
    for img in (128, 512, 1024):
        assert schedulers.resolve("flux", 50, img)[0] == O.resolve_flux_sigma(50, img)
    sig, a, b, s = schedulers.resolve("flux", 50, 1024)
    assert abs(sig - 0.19545) < 1e-4 and abs(a + b - 1.0) < 1e-7 and s == 1.0
    with pytest.raises(ValueError):
        schedulers.resolve("flux", 0, 1024)          # strength 0 leaves no step (pipeline_flux_img2img.py:768-772)


def test_oracle_feature_resize_matches_reference_store():
    """tests/golden/feature_store_resize.pt: the reference's real FeatureStore.store with resize_ratio 2 and 3
    (feature_extractor.py:51-53, adaptive_avg_pool2d; 3 does not divide the 16 x 16 maps)."""
    gold = torch.load(os.path.join(GOLD, "feature_store_resize.pt"), weights_only=False)
    for r in (2, 3):
        st = O.FeatureStore({"a": True, "b": True}, r)
        st.store(gold["conv"], "a")
        st.store(gold["vit"], "b")
        for k in ("a", "b"):
            assert st.feats[k].shape == gold["r%d" % r][k].shape == (2, gold["conv"].shape[1] if k == "a" else 128, 16 // r, 16 // r)
            assert torch.allclose(st.feats[k], gold["r%d" % r][k], atol=1e-6)


@pytest.mark.parametrize("mode", ["default", "sample_first_original", "aggregate_nested"])
def test_cli_on_disk_format_matches_reference(mode, tmp_path):
    """tests/golden/cli_layouts.json holds the file trees (paths, shapes, dtypes, content hashes) the reference's OWN
    extract_feature.py main() wrote for three flag combinations on a stand-in extractor (tools/make_golden.py);
    generic_diffusion_feature_b200.extract_feature must write byte-identical arrays at identical paths."""
    from common import CLI_MODES, FakeExtractor, cli_fixture_inputs, tree_digest
    from generic_diffusion_feature_b200 import extract_feature as cli
    gold = json.load(open(os.path.join(GOLD, "cli_layouts.json")))[mode]
    root = str(tmp_path)
    cli_fixture_inputs(root)
    od = os.path.join(root, "out")
    argv = ["--layer", "x.json", "--t", "50", "-b", "4", "--input_dir", os.path.join(root, "imgs", "*", "*.png"),
            "--prompt_file", os.path.join(root, "prompt.txt"), "--output_dir", od, "--writer_threads", "3"] + CLI_MODES[mode]
    n = cli.run(cli.build_parser().parse_args(argv), extractor=FakeExtractor())
    assert n == 6
    assert tree_digest(od) == gold


def test_cli_nearest_resize_matches_torch_interpolate():
    """The aggregated CLI output resizes with F.interpolate's default mode (`nearest`, extract_feature.py:122-124);
    the writer threads do it in numpy - same index rule for integer and non-integer ratios, up and down."""
    import torch.nn.functional as F
    from generic_diffusion_feature_b200.extract_feature import nearest_resize_chw
    g = torch.Generator().manual_seed(5)
    for h, size in ((4, 16), (8, 8), (6, 16), (16, 6), (5, 7), (32, 128), (3, 128)):
        a = torch.randn(3, h, h, generator=g)
        want = F.interpolate(a[None], size)[0].numpy()
        got = nearest_resize_chw(a.numpy(), size)
        assert got.shape == want.shape and np.array_equal(got, want), (h, size)


def test_cli_two_rank_shards_write_the_same_tree(tmp_path, monkeypatch):
    """Two ranks (RANK / WORLD_SIZE as torchrun sets them) each extract their contiguous shard into one output tree:
    together they produce exactly the single-process tree of the reference (global indices name the files)."""
    from common import CLI_MODES, FakeExtractor, cli_fixture_inputs, tree_digest
    from generic_diffusion_feature_b200 import extract_feature as cli
    gold = json.load(open(os.path.join(GOLD, "cli_layouts.json")))["default"]
    root = str(tmp_path)
    cli_fixture_inputs(root)
    od = os.path.join(root, "out")
    argv = ["--layer", "x.json", "--t", "50", "-b", "2", "--input_dir", os.path.join(root, "imgs", "*", "*.png"),
            "--prompt_file", os.path.join(root, "prompt.txt"), "--output_dir", od, "--device", "cpu"] + CLI_MODES["default"]
    counts = []
    for rank in range(2):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("WORLD_SIZE", "2")
        counts.append(cli.run(cli.build_parser().parse_args(argv), extractor=FakeExtractor()))
    assert counts == [3, 3]
    assert tree_digest(od) == gold


def test_dit_param_specs_and_pos_embed_match_oracle():
    m = _models()
    for ver in ("pixart-sigma", "pixart-sigma-512"):
        cfg = m.DIT_CONFIGS[ver]
        assert cfg == O.DIT_CONFIGS[ver]
        with torch.device("meta"):
            model = O.PixArtTransformer2DModel(cfg)
        want = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        assert dict(m.dit_param_specs(cfg)) == want
    n = sum(int(np.prod(s)) for k, s in m.dit_param_specs(m.DIT_CONFIGS["pixart-sigma"]) if "pos_embed.pos_embed" not in k)
    assert 0.60e9 < n < 0.62e9              # PixArt-Sigma-XL/2: 0.61 B parameters
    a = m.sincos_pos_embed_2d(64, 8, 8, 2.0)
    b = O.sincos_pos_embed_2d(64, 8, 8, 2.0)
    assert a.shape == (64, 64) and (a - b).abs().max().item() < 1e-6
    # first half encodes x (fastest-varying token index), second half y
    assert (a[0, :32] - a[8, :32]).abs().max().item() < 1e-7 and (a[0, 32:] - a[1, 32:]).abs().max().item() < 1e-7


def test_oracle_correspondence_matches_reference():
    gold = torch.load(os.path.join(GOLD, "correspondence.pt"), weights_only=False)
    pts = gold["points"].numpy()
    p2, _ = O.find_nn_source_correspondences(gold["f1"], gold["f2"], pts, tuple(gold["load_size"]))
    assert torch.equal(p2, gold["points2"])
    assert np.array_equal(O.points_to_idxs(pts, tuple(gold["load_size"])), gold["idx"].numpy())


def test_oracle_segmentor_head_matches_reference():
    """tests/golden/segmentor_head.pt: outputs of the reference's real ResBlock / MultiRes / DiffusionSegmentor.extract_feat
    (segmentation/models/diffusion_segmentor.py, eval mode; tools/make_golden.py) vs the oracle restatement, and the
    BatchNorm fold the CUDA head applies at load time vs the unfolded form."""
    gold = torch.load(os.path.join(GOLD, "segmentor_head.pt"), weights_only=False)
    sd = {k: v.float() for k, v in gold["state_dict"].items()}
    outs = O.seg_extract_feat(gold["features"], gold["feature_layers"], sd)
    for got, want in zip(outs, gold["outs"]):
        assert got.shape == want.shape
        assert (got - want.float()).abs().max().item() <= 2e-3 * want.float().abs().max().item()   # fixture is fp16-rounded
    x = gold["features"]["up-level1-upsampler-out"].float()
    y = x
    for _ in range(gold["multires_n"]):
        y = O.seg_resblock(y, sd, "up_level1_upsampler_out")
    assert (y - gold["multires_out"].float()).abs().max().item() <= 2e-3 * y.abs().max().item()
    # several-extractors branch (diffusion_segmentor.py:248-297): MultiRes per map / per level sum, 'amalgemated' ResBlock
    mg = gold["multi"]
    msd = {k: v.float() for k, v in mg["state_dict"].items()}
    mouts = O.seg_extract_feat_multi(mg["features"], mg["feature_layers"], mg["c_per_level"], msd)
    for got, want in zip(mouts, mg["outs"]):
        assert got.shape == want.shape
        assert (got - want.float()).abs().max().item() <= 2e-3 * want.float().abs().max().item()
    # fold: conv(x, w * s) + (b - mean) * s + beta == BN(conv(x, w) + b)
    p = "up_level1_upsampler_out.conv1"
    s = sd[p + ".1.weight"] * torch.rsqrt(sd[p + ".1.running_var"] + 1e-5)
    folded = F.conv2d(x, sd[p + ".0.weight"] * s[:, None, None, None], (sd[p + ".0.bias"] - sd[p + ".1.running_mean"]) * s
                      + sd[p + ".1.bias"], padding=1)
    plain = F.batch_norm(F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], padding=1), sd[p + ".1.running_mean"],
                         sd[p + ".1.running_var"], sd[p + ".1.weight"], sd[p + ".1.bias"], False, 0.0, 1e-5)
    assert torch.allclose(folded, plain, atol=1e-4, rtol=1e-4)


def test_vae_out_host_logic_matches_oracle():
    """`vae-out` host side (diffusion_feature.py:477-485): scheduler.step coefficients of the product equal the
    oracle's for every UNet family over the timestep range, the Euler step equals the textbook form computed from the
    sigma table, and the decoder parameter list equals the oracle decoder's state dict (names + shapes)."""
    from generic_diffusion_feature_b200 import schedulers
    models = _models()
    for version in ("xl", "pgv2", "2-1", "1-5"):
        for t in (1, 2, 50, 261, 500, 998, 999):
            a, b = schedulers.step_coeffs(version, t), O.scheduler_step_coeffs(version, t)
            assert abs(a[0] - b[0]) < 1e-6 and abs(a[1] - b[1]) < 1e-6, (version, t, a, b)   # fp32 cumprod tables: numpy vs torch
    # Euler: prev = x + eps * (sigma_next - sigma); at t = 50 (timestep 50 of [1000..1]) sigma_next = sigma(49)
    ac = schedulers.alphas_cumprod().astype(np.float64)
    sig = lambda k: ((1 - ac[k]) / ac[k]) ** 0.5
    c_s, c_m = schedulers.step_coeffs("xl", 50)
    assert c_s == 1.0 and abs(c_m - (sig(49) - sig(50))) < 1e-12
    # PNDM first PLMS step reproduces x_{t-1} of the DDIM-like closed form when noise_pred is the true noise:
    # x_t = sqrt(a_t) z + sqrt(1 - a_t) eps  ->  _get_prev_sample gives approximately sqrt(a_p) z + sqrt(1 - a_p) eps
    ts = int(schedulers.resolve("1-5", 261)[0])
    c_s, c_m = schedulers.step_coeffs("1-5", 261)
    z_coef = c_s * ac[ts] ** 0.5
    e_coef = c_s * (1 - ac[ts]) ** 0.5 + c_m
    assert abs(z_coef - ac[ts - 1] ** 0.5) < 1e-9 and abs(e_coef - (1 - ac[ts - 1]) ** 0.5) < 2e-4
    with pytest.raises(NotImplementedError):
        schedulers.step_coeffs("pixart-sigma", 50)
    vae = O.Vae(TINY_VAE["scaling_factor"], decoder=True, block_out=TINY_VAE["block_out"], layers=TINY_VAE["layers"],
                latent=TINY_VAE["latent"], eps=TINY_VAE["eps"])
    want = {k: tuple(v.shape) for k, v in vae.state_dict().items() if k.startswith(("decoder.", "post_quant_conv."))}
    got = {n: tuple(s) for n, s in models.vae_decoder_param_specs(TINY_VAE)}
    assert got == want
    sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE, with_decoder=True)
    assert "vae.decoder.conv_out.weight" in sd and "vae.decoder.conv_in.weight" not in models.synthetic_state_dict(
        "xl", "cpu", TINY_XL, TINY_VAE)


def test_oracle_whole_path_digest():
    gold = torch.load(os.path.join(GOLD, "extract_tiny_xl.pt"), weights_only=False)
    sd = _models().synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    unet, vae = build_oracle(TINY_XL, TINY_VAE, sd)
    store = O.FeatureStore({i: True for i in gold["ids"]})
    O.attach_gatherers(unet, store)
    image, ctx, pooled, ev, eq = make_inputs(1, 64, TINY_XL["ctx_dim"], 64)
    feats, latents, npred = O.extract("xl", unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=64)
    assert (latents - gold["latents"]).abs().max().item() < 1e-4
    assert (npred - gold["noise_pred"]).abs().max().item() < 1e-3
    for k, (m, s, mx) in gold["stats"].items():
        assert abs(float(feats[k].std()) - s) <= 1e-3 * max(1.0, s), k
    for k, v in gold["feats_subset"].items():
        assert (feats[k] - v.float()).abs().max().item() <= 2e-3 * max(1.0, v.float().abs().max().item()), k


# ------------------------------------------------------------------------------------------ host logic
def test_id_grammar_matches_reference_configs():
    """The planner's id list == the non-map keys of feature/configs/config_{xl,15}_full.json, same order."""
    from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
    ref = json.load(open(os.path.join(GOLD, "reference_ids.json")))
    m = _models()
    assert _unet_feature_ids(m.UNET_CONFIGS["xl"]) == ref["ids_xl"] and len(ref["ids_xl"]) == 472
    assert _unet_feature_ids(m.UNET_CONFIGS["1-5"]) == ref["ids_1-5"] and len(ref["ids_1-5"]) == 165
    # with the attention-probability maps interleaved (what accept-all plans, like the reference's empty-config mode that
    # produced these files): all 612 / 197 keys of config_{xl,15}_full.json in file (= execution) order
    assert _unet_feature_ids(m.UNET_CONFIGS["xl"], with_maps=True) == ref["all_ids_xl"] and len(ref["all_ids_xl"]) == 612
    assert _unet_feature_ids(m.UNET_CONFIGS["1-5"], with_maps=True) == ref["all_ids_1-5"] and len(ref["all_ids_1-5"]) == 197
    # every practical / legacy config id is a legal id of its architecture
    for f, ver in (("config_xl_practical.json", "xl"), ("config_xl_legacy.json", "xl"),
                   ("config_15_practical.json", "1-5"), ("config_15_legacy.json", "1-5")):
        legal = set(_unet_feature_ids(m.UNET_CONFIGS[ver]))
        assert set(ref[f].keys()) <= legal, f


def _channels_of(fid, cfg):
    """Channel count of a feature id from the architecture alone (res/vit -> level width, ffn-inner -> 4x)."""
    bo = cfg["block_out"]
    n = len(bo)
    if fid.startswith("mid"):
        c = bo[-1]
    elif fid.startswith("down"):
        c = bo[int(re.match(r"down-level(\d)", fid).group(1))]
    elif fid.startswith("up"):
        c = bo[n - 1 - int(re.match(r"up-level(\d)", fid).group(1))]
    else:
        c = 4 if fid in ("unet-in", "unet-out") else bo[0]
    return 4 * c if "ffn-inner" in fid else c


def test_feature_len_of_reference_task_configs():
    """correspondence/correspondence/config_*.json:feature_len == channel sum of the layer config they name."""
    ref = json.load(open(os.path.join(GOLD, "reference_ids.json")))
    m = _models()
    for corr, layer, ver in (("corr_config_sdxl.json", "config_xl_practical.json", "xl"),
                             ("corr_config_legacy_sdxl.json", "config_xl_legacy.json", "xl"),
                             ("corr_config_sd15.json", "config_15_practical.json", "1-5"),
                             ("corr_config_legacy_sd15.json", "config_15_legacy.json", "1-5")):
        total = sum(_channels_of(k, m.UNET_CONFIGS[ver]) for k, v in ref[layer].items() if v)
        assert total == ref[corr]["feature_len"], (corr, total)


@pytest.mark.parametrize("version", ["xl", "2-1", "1-5"])
def test_param_specs_match_oracle_modules(version):
    """Parameter names / shapes of the packer == state_dict of the oracle modules (diffusers naming)."""
    m = _models()
    with torch.device("meta"):
        unet = O.UNet2DConditionModel(O.UNET_CONFIGS[version])
        vcfg = m.VAE_CONFIGS[version]
        vae = O.Vae(vcfg["scaling_factor"], block_out=vcfg["block_out"], layers=vcfg["layers"], latent=vcfg["latent"],
                    eps=vcfg["eps"])
    want = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
    got = dict(m.unet_param_specs(m.UNET_CONFIGS[version]))
    assert got == want
    wantv = {k: tuple(v.shape) for k, v in vae.state_dict().items()}
    assert dict(m.vae_param_specs(vcfg)) == wantv
    if version == "xl":
        n = sum(int(np.prod(s)) for s in got.values())
        assert 2.55e9 < n < 2.60e9           # SDXL UNet, 2.57 B parameters


def test_synthetic_weights_are_deterministic_by_name():
    m = _models()
    a = m.init_param("unet.conv_in.weight", (320, 4, 3, 3))
    b = m.init_param("unet.conv_in.weight", (320, 4, 3, 3))
    c = m.init_param("unet.conv_out.weight", (4, 320, 3, 3))
    assert torch.equal(a, b) and a.shape != c.shape
    assert abs(float(m.init_param("x.norm1.weight", (4096,)).mean()) - 1.0) < 0.01


@pytest.mark.parametrize("version,t,want_ts", [("xl", 50, 50.0), ("2-1", 50, 49.0), ("1-5", 50, 51.0),
                                               ("xl", 250, 250.0), ("1-5", 1, 2.0), ("pixart-sigma", 50, 50.0),
                                               ("pixart-sigma-512", 261, 261.0)])
def test_scheduler_resolution(version, t, want_ts):
    from generic_diffusion_feature_b200 import schedulers
    ts, a, b, s = schedulers.resolve(version, t)
    ots, oa, ob, os_ = O.resolve_timestep(version, t)
    assert ts == want_ts == ots
    assert abs(a - oa) < 1e-6 and abs(b - ob) < 1e-5 and abs(s - os_) < 1e-6
    if version.startswith("pixart"):
        assert s == 1.0 and abs(a * a + b * b - 1.0) < 1e-5     # variance-preserving q_sample, linear betas
        return
    # both forms are the same q_sample: (a*z + b*eps)*s = sqrt(abar) z + sqrt(1-abar) eps
    abar = float(O.alphas_cumprod()[int(ts)])
    assert abs(a * s - abar ** 0.5) < 1e-5 and abs(b * s - (1 - abar) ** 0.5) < 1e-5


def test_error_behaviour_mirrors_reference():
    m = _models()
    with pytest.raises(NotImplementedError):
        m.get_diffusion_model("xl", "bfloat16")          # models.py:15-16
    with pytest.raises(NotImplementedError):
        m.get_diffusion_model("sd-3", "float16")         # models.py:173-174
    from generic_diffusion_feature_b200.components.feature_extractor import FeatureStore, prepare_feature_extractor
    fs = prepare_feature_extractor("xl", None, {"a": True, "b": False}, 1, False)
    assert isinstance(fs, FeatureStore) and not fs.accept_all and fs.to_store == {"a": True, "b": False}
    assert prepare_feature_extractor("xl", None, None, 1, False).accept_all      # feature_extractor.py:10-15
    first = fs.stored_feats
    fs.reset()
    assert fs.stored_feats is not first                  # reset() rebinds (feature_extractor.py:28-29)


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of falling back."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from generic_diffusion_feature_b200._lib import GdfError
    m = _models()
    with pytest.raises(GdfError):
        m.B200Pipe("xl", m.UNET_CONFIGS["xl"], m.VAE_CONFIGS["xl"], "cpu")
    # the downstream heads have no CPU path either: building one without a GPU fails loudly
    from generic_diffusion_feature_b200 import segmentation as S
    gold = torch.load(os.path.join(GOLD, "segmentor_head.pt"), weights_only=False)
    with pytest.raises(Exception):
        S.SegmentorFeatureHead(gold["feature_layers"], {k: v.float() for k, v in gold["state_dict"].items()})
    assert S.layer_conv_name("up-level0-upsampler-out") == "up_level0_upsampler_out"      # diffusion_segmentor.py:188-192
    assert S.layer_conv_name("sum1", 0) == "0_sum1"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "generic_diffusion_feature_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                imports = [l for l in src.splitlines() if re.match(r"\s*(from|import)\s", l)]
                assert not any("oracle" in l for l in imports), "%s imports the oracle" % f


# ------------------------------------------------------------------------------------------ C ABI surface
def test_abi_exports_every_declared_symbol():
    from generic_diffusion_feature_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from generic_diffusion_feature_b200 import build
        build.build(verbose=False)
    header = open(os.path.join(ROOT, "include", "gdf.h")).read()
    declared = sorted(set(re.findall(r"\b(gdf_[a-z0-9_]+)\s*\(", header)))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "declared in gdf.h but not exported: %s" % missing
    assert sorted(_lib.EXPORTS) == declared, "python binding list out of sync with gdf.h"
    lib.gdf_abi_version.restype = ctypes.c_int
    assert lib.gdf_abi_version() == 3
    # error plumbing without touching the GPU: null handle -> negative code + message
    lib.gdf_last_error.restype = ctypes.c_char_p
    assert lib.gdf_plan(None, None, 0, 1, 1024, None, None) < 0
    assert b"null handle" in lib.gdf_last_error()


def test_struct_layouts_match_header():
    """ctypes mirrors of the C structs have the sizes the C compiler produces (checked with a tiny gcc probe)."""
    import subprocess
    import tempfile
    from generic_diffusion_feature_b200 import _lib
    src = ('#include <stdio.h>\n#include "gdf.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gdf_epilogue),'
           'sizeof(gdf_capture_seg), sizeof(gdf_unet_arch), sizeof(gdf_vae_arch), sizeof(gdf_slot),'
           'sizeof(gdf_resize_src), sizeof(gdf_dit_arch), sizeof(gdf_flux_arch));return 0;}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o",
                               os.path.join(d, "p")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "p")]).split()]
    got = [ctypes.sizeof(c) for c in (_lib.Epilogue, _lib.CaptureSeg, _lib.UNetArch, _lib.VaeArch, _lib.Slot,
                                      _lib.ResizeSrc, _lib.DitArch, _lib.FluxArch)]
    assert got == sizes


# ------------------------------------------------------------------------------------------ sharding (gloo, 2 ranks)
def _worker(rank, world, port, q):
    import torch.distributed as dist
    from generic_diffusion_feature_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n_pairs = 7
    s, e = parallel.shard_range(n_pairs, rank, world)
    local = torch.arange(s, e, dtype=torch.int64)[:, None] * torch.ones(1, 3, dtype=torch.int64)
    counts = [parallel.shard_range(n_pairs, r, world)[1] - parallel.shard_range(n_pairs, r, world)[0]
              for r in range(world)]
    allr = parallel.gather_to_rank0(local, counts)
    mx = parallel.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((allr[:, 0].tolist(), mx))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    import torch.multiprocessing as mp
    from generic_diffusion_feature_b200 import parallel
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rows == list(range(7)) and mx == 11.0


def test_checkpoint_dir_round_trip(tmp_path):
    """get_diffusion_model's checkpoint source (reference: Pipeline.from_pretrained, models.py:18-172): a diffusers-layout
    directory written by save_diffusers_dir reads back name-for-name and bit-for-bit through load_diffusers_dir; VAE
    checkpoints in the pre-0.18 naming (query / key / value / proj_attn stored as 1x1 convs) are converted; decoder
    tensors are dropped."""
    m = _models()
    from common import TINY_VAE, TINY_XL
    sd = m.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
    d = str(tmp_path / "ckpt")
    m.save_diffusers_dir(sd, d)
    back = m.load_diffusers_dir(d, "xl")
    assert set(back) == set(sd)
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    assert set(back) == set(m.expected_shapes(TINY_XL, TINY_VAE))
    # old VAE naming + a decoder tensor
    legacy = {}
    for k, v in sd.items():
        if not k.startswith("vae."):
            continue
        k2 = k[4:]
        for new, old in ((".to_q.", ".query."), (".to_k.", ".key."), (".to_v.", ".value."), (".to_out.0.", ".proj_attn.")):
            if new in k2:
                k2 = k2.replace(new, old)
                if k2.endswith(".weight"):
                    v = v[:, :, None, None]
        legacy[k2] = v.contiguous()
    legacy["decoder.conv_in.weight"] = torch.zeros(4, 4, 3, 3)
    from safetensors.torch import save_file
    save_file(legacy, os.path.join(d, "vae", "diffusion_pytorch_model.safetensors"))
    back = m.load_diffusers_dir(d, "xl")
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # with the decoder (vae-out): its tensors round-trip too (mid-block attention in the old naming included) and are
    # dropped unless asked for
    sd_dec = m.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE, with_decoder=True)
    d2 = str(tmp_path / "ckpt_dec")
    m.save_diffusers_dir(sd_dec, d2)
    assert set(m.load_diffusers_dir(d2, "xl")) == set(sd)
    back = m.load_diffusers_dir(d2, "xl", with_decoder=True)
    assert set(back) == set(sd_dec) and all(torch.equal(back[k], sd_dec[k]) for k in sd_dec)
    assert set(back) - set(sd) == set(m.optional_shapes(TINY_VAE))


def test_no_silent_synthetic_weights(monkeypatch):
    """Without state_dict / model_dir / GDF_MODEL_DIR / an explicit synthetic opt-in the factory raises (the reference
    would download a checkpoint here; a random network must never be a silent default). Checked before any device use."""
    m = _models()
    from generic_diffusion_feature_b200._lib import GdfError
    monkeypatch.delenv("GDF_MODEL_DIR", raising=False)
    monkeypatch.delenv("GDF_SYNTHETIC", raising=False)
    with pytest.raises(GdfError, match="no weights"):
        m.get_diffusion_model("xl", "float16", device="cuda:0")
    with pytest.raises(FileNotFoundError):
        m.get_diffusion_model("xl", "float16", device="cuda:0", model_dir="/nonexistent/gdf-model-dir")
    assert "pixart-alpha" in m.DIT_CONFIGS and m.VAE_CONFIGS["pixart-alpha"]["scaling_factor"] == 0.18215


def test_cli_prefetch_error_reaches_the_consumer(tmp_path):
    """A corrupt / missing input file must raise in the consumer, not leave it blocked on the queue."""
    from generic_diffusion_feature_b200.extract_feature import prefetch_images
    from PIL import Image
    good = str(tmp_path / "a.png")
    Image.new("RGB", (8, 8)).save(good)
    bad = str(tmp_path / "b.png")
    open(bad, "wb").write(b"not an image")
    got = []
    with pytest.raises(RuntimeError, match="image prefetch failed"):
        for first, imgs in prefetch_images([(good, "a"), (bad, "b")], 1, 8):
            got.append(first)
    assert got == [0]
