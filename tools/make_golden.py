"""Generate the golden fixtures under tests/golden/ by EXECUTING the reference's own code in this container.

    python tools/make_golden.py            (needs /root/reference; cannot run on the GPU box)

Fixtures (small, committed; regenerate with this script):
  unet_tiny_xl.pt / unet_tiny_21.pt
      Output of the reference's vendored UNet2DConditionModel.forward (feature/diffusers/models/unet/
      unet_2d_condition.py) + the reference's real prepare_feature_extractor / FeatureStore
      (feature/components/feature_extractor.py, train_unet=True so nothing is cast) on the reduced-width
      topologies of tests/common.py, synthetic weights by name, seeded inputs. Stored: every feature map (fp16),
      the noise prediction, and the inputs. The oracle and the CUDA path are both checked against these.
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
from common import O, TINY_21, TINY_VAE, TINY_XL, build_oracle, make_inputs  # noqa: E402
from generic_diffusion_feature_b200.components import models  # noqa: E402
from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids  # noqa: E402

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
    golden_correspondence()
    golden_extract()
