"""vae-out at SDXL size (1024^2): time of the decoder pass on top of the normal forward, finite check, per-kind split."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_inputs  # noqa: E402
from generic_diffusion_feature_b200.components import models  # noqa: E402
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd = models.synthetic_state_dict("xl", "cuda:0", with_decoder=True)
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd)
del sd
image, ctx, pooled, ev, eq = make_inputs(B, 1024, 2048, 1280)
res = {}
for name, layer in (("plain", {"unet-out": True}), ("vae-out", {"unet-out": True, "vae-out": True})):
    fe = FeatureExtractor(layer, "xl", "cuda:0", img_size=1024, external_model=pipe)
    args = ((ctx, ctx, pooled, pooled), B, image.cuda())
    kw = dict(image_type="tensors", t=50, noise=(ev, eq))
    for _ in range(2):
        got = fe.extract(*args, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        got = fe.extract(*args, **kw)
    e1.record()
    torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 3
    if name == "vae-out":
        v = got["vae-out"]
        print("vae-out", tuple(v.shape), v.dtype, "finite", bool(torch.isfinite(v.float()).all()), "absmax %.3f" % v.float().abs().max().item())
print("B = %d: forward %.1f ms, forward + vae-out %.1f ms -> decoder pass %.1f ms (%.2f ms / image)"
      % (B, res["plain"], res["vae-out"], res["vae-out"] - res["plain"], (res["vae-out"] - res["plain"]) / B))
print("workspace GB", pipe.lib.gdf_workspace_bytes(pipe.handle) / 1e9 if hasattr(pipe.lib, "gdf_workspace_bytes") else None)
