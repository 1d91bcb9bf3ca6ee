import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from common import TINY_XL, TINY_VAE, make_inputs
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor
sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd, unet_cfg=TINY_XL, vae_cfg=TINY_VAE)
image, ctx, pooled, ev, eq = make_inputs(2, 128, TINY_XL["ctx_dim"], 64)
ids = _unet_feature_ids(TINY_XL)
fe = FeatureExtractor({i: True for i in ids}, "xl", "cuda:0", img_size=128, external_model=pipe)
img = image.cuda()
run = lambda: {k: v.clone() for k, v in fe.extract((ctx, ctx, pooled, pooled), 2, img, image_type="tensors",
                                                   t=50, noise=(ev, eq)).items()}
outs = [run() for _ in range(5)]
torch.cuda.synchronize()
for i, o in enumerate(outs[:2]):
    bad = [k for k in ids if not torch.isfinite(o[k].float()).all()]
    print("run %d: %d maps with non-finite values; first %s" % (i, len(bad), bad[:3]), flush=True)
for i in range(4):
    diff = [k for k in ids if not torch.equal(outs[i][k], outs[i + 1][k])]
    d0 = diff[0] if diff else None
    msg = ""
    if d0:
        x, y = outs[i][d0].float(), outs[i + 1][d0].float()
        nz = (x != y)
        msg = " first %s: %d of %d elements differ, max |d| %.3e, where (first) %s" % (d0, int(nz.sum()), nz.numel(), float((x - y).abs().max()), nz.nonzero()[:3].tolist())
    print("run %d vs %d: %d of %d maps differ%s" % (i, i + 1, len(diff), len(ids), msg), flush=True)
