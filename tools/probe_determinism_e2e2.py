"""Does a result depend on what the PREVIOUS call computed (stale-buffer read) or only on its own inputs?
Sequence A A B A B B A with two different images / noises; all A results must be bit-identical, all B results too."""
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
g = torch.Generator().manual_seed(99)
imageB = torch.rand(image.shape, generator=g) * 2 - 1
evB, eqB = torch.randn(ev.shape, generator=g), torch.randn(eq.shape, generator=g)
ids = _unet_feature_ids(TINY_XL)
fe = FeatureExtractor({i: True for i in ids}, "xl", "cuda:0", img_size=128, external_model=pipe)
if os.environ.get("PROBE_PREPLAN") == "1":
    fe._ensure_plan(2, 77); torch.cuda.synchronize()
def run(which):
    im, a, b = (image, ev, eq) if which == "A" else (imageB, evB, eqB)
    out = fe.extract((ctx, ctx, pooled, pooled), 2, im.cuda(), image_type="tensors", t=50, noise=(a, b))
    r = {k: v.clone() for k, v in out.items()}
    torch.cuda.synchronize()
    return r
seq = "AABABBA"
outs = [run(c) for c in seq]
ref = {}
for i, c in enumerate(seq):
    if c not in ref:
        ref[c] = i
        print("call %d (%s): reference for %s" % (i, c, c)); continue
    diff = [k for k in ids if not torch.equal(outs[i][k], outs[ref[c]][k])]
    print("call %d (%s, after %s): %d of %d maps differ from call %d%s" % (i, c, seq[i - 1], len(diff), len(ids), ref[c], (" first " + diff[0]) if diff else ""), flush=True)
