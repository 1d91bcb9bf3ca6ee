"""VAE encode right after plan / after another encode / after a UNet forward: are the latents bit-identical?"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from common import TINY_XL, TINY_VAE, make_inputs
from generic_diffusion_feature_b200 import _lib
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor
sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd, unet_cfg=TINY_XL, vae_cfg=TINY_VAE)
image, ctx, pooled, ev, eq = make_inputs(2, 128, TINY_XL["ctx_dim"], 64)
ids = _unet_feature_ids(TINY_XL)
fe = FeatureExtractor({i: True for i in ids}, "xl", "cuda:0", img_size=128, external_model=pipe)
fe._ensure_plan(2, 77)
lib = pipe.lib
img, evd, eqd = image.cuda().float().contiguous(), ev.cuda().float().contiguous(), eq.cuda().float().contiguous()
def enc():
    lat = torch.zeros(2, 4, 16, 16, device="cuda")
    _lib.check(lib.gdf_encode_noise(pipe.handle, _lib.ptr(img), _lib.ptr(evd), _lib.ptr(eqd), 1.0, 0.5, 1.0, _lib.ptr(lat), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return lat.clone()
def full():
    out = fe.extract((ctx, ctx, pooled, pooled), 2, img, image_type="tensors", t=50, noise=(ev, eq))
    torch.cuda.synchronize()
    return out["unet-in"].clone()
a = enc(); b = enc()
u0 = full()
c = enc(); d = enc()
u1 = full(); u2 = full()
e = enc()
n = lambda x, y: int((x != y).sum())
print(os.environ.get("PROBE_TAG", ""), "enc,enc: %d | enc after UNet vs first: %d | enc after that vs first: %d | unet-in call0 vs call1: %d, call1 vs call2: %d | last enc vs first: %d"
      % (n(a, b), n(a, c), n(a, d), n(u0, u1), n(u1, u2), n(a, e)), flush=True)
