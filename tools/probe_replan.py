"""GPU probe: run-to-run and re-plan differences of every map of the tiny SDXL topology, vs the oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch, torch.nn.functional as F
from common import O, TINY_XL, TINY_VAE, build_oracle, make_inputs
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor

sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd, unet_cfg=TINY_XL, vae_cfg=TINY_VAE)
image, ctx, pooled, ev, eq = make_inputs(2, 128, TINY_XL["ctx_dim"], 64)
ids = _unet_feature_ids(TINY_XL)
layer = {i: True for i in ids}
unet, vae = build_oracle(TINY_XL, TINY_VAE, sd)
store = O.FeatureStore(layer); O.attach_gatherers(unet, store)
want, _, _ = O.extract("xl", unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=128)
fe_a = FeatureExtractor(layer, "xl", "cuda:0", img_size=128, external_model=pipe)
fe_b = FeatureExtractor({"mid-vit-out": True, "unet-out": True}, "xl", "cuda:0", img_size=128, external_model=pipe)
run = lambda fe: {k: v.float().cpu() for k, v in fe.extract((ctx, ctx, pooled, pooled), 2, image.cuda(),
                                                            image_type="tensors", t=50, noise=(ev, eq)).items()}
a1 = run(fe_a); a1b = run(fe_a); b1 = run(fe_b); a2 = run(fe_a); a3 = run(fe_a)
def st(x, y):
    return "%.6f/%.4f" % (F.cosine_similarity(x.flatten(), y.flatten(), dim=0).item(),
                          (x - y).abs().max().item() / max(1e-6, y.abs().max().item()))
print("id : a1~a1b  a1~a2  a2~a3 | a1~oracle a2~oracle   (cos/maxrel)")
for k in ids:
    w = want[k].float()
    row = (st(a1[k], a1b[k]), st(a1[k], a2[k]), st(a2[k], a3[k]), st(a1[k], w), st(a2[k], w))
    flag = "" if all(float(r.split("/")[0]) > 0.99995 for r in row[:3]) else "  <<<"
    if flag or "ffn" in k or k.endswith("vit-out"):
        print(k, *row, flag)
