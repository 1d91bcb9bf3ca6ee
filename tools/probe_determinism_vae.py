"""Is the VAE-encode + q_sample part (gdf_encode_noise) bit-reproducible from its first call on? Prints per-call equality."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from common import TINY_XL, TINY_VAE, make_inputs
from generic_diffusion_feature_b200 import _lib
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor
img_size = int(os.environ.get("PROBE_IMG", "128"))
sd = models.synthetic_state_dict("xl", "cpu", TINY_XL, TINY_VAE)
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd, unet_cfg=TINY_XL, vae_cfg=TINY_VAE)
image, ctx, pooled, ev, eq = make_inputs(2, img_size, TINY_XL["ctx_dim"], 64)
fe = FeatureExtractor({"unet-out": True}, "xl", "cuda:0", img_size=img_size, external_model=pipe)
fe._ensure_plan(2, 77)
lib = pipe.lib
img, evd, eqd = image.cuda().float().contiguous(), ev.cuda().float().contiguous(), eq.cuda().float().contiguous()
L = img_size // 8
outs = []
for i in range(5):
    lat = torch.zeros(2, 4, L, L, device="cuda")
    _lib.check(lib.gdf_encode_noise(pipe.handle, _lib.ptr(img), _lib.ptr(evd), _lib.ptr(eqd), 1.0, 0.5, 1.0, _lib.ptr(lat), _lib.stream_ptr()))
    torch.cuda.synchronize()
    outs.append(lat.clone())
print(os.environ.get("PROBE_TAG", ""), [int((outs[i] != outs[i + 1]).sum()) for i in range(4)], "elements differ between consecutive calls (of %d)" % outs[0].numel(), flush=True)
