"""Attention-probability maps at SDXL size: the tensor-core form (batched tcgen05 GEMMs + softmax) against the CUDA-core
kernel (GDF_MAPS_TC=0). Run once per setting; the second run compares with the maps the first one saved.

    python tools/probe_maps_tc.py save /tmp/maps_tc.pt ; GDF_MAPS_TC=0 python tools/probe_maps_tc.py cmp /tmp/maps_tc.pt
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_inputs  # noqa: E402
from generic_diffusion_feature_b200.components import models  # noqa: E402
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor  # noqa: E402

mode, path = sys.argv[1], sys.argv[2]
ids = ["down-level1-repeat0-vit-block0-self-map", "down-level2-repeat0-vit-block0-self-map",
       "down-level2-repeat0-vit-block0-cross-map", "down-level1-repeat0-vit-block0-out", "down-level2-repeat0-vit-block0-out",
       "up-level0-repeat0-vit-block3-self-map", "unet-out"]
sd = models.synthetic_state_dict("xl", "cuda:0")
pipe = models.get_diffusion_model("xl", "float16", device="cuda:0", state_dict=sd)
del sd
B = 2
image, ctx, pooled, ev, eq = make_inputs(B, 1024, 2048, 1280)
fe = FeatureExtractor({i: True for i in ids}, "xl", "cuda:0", img_size=1024, external_model=pipe)
args = ((ctx, ctx, pooled, pooled), B, image.cuda())
kw = dict(image_type="tensors", t=50, noise=(ev, eq))
got = fe.extract(*args, **kw)
torch.cuda.synchronize()
lib, h = pipe.lib, pipe.handle
lib.gdf_profile(h, 1)
got = fe.extract(*args, **kw)
torch.cuda.synchronize()
lib.gdf_profile(h, 0)
csv = "/tmp/maps_perop.csv"
lib.gdf_profile_dump(h, csv.encode())
for line in open(csv):
    if "attention-probs" in line:
        print(line.strip()[:200])
t0 = time.time()
for _ in range(3):
    got = fe.extract(*args, **kw)
torch.cuda.synchronize()
print("3 forwards with maps: %.1f ms each (B = %d)" % ((time.time() - t0) / 3 * 1e3, B))
m = got[ids[0]]
print("map", tuple(m.shape), m.dtype, "row sums - 1: max %.2e" % (m.float().sum(-1) - 1).abs().max().item())
if mode == "save":
    torch.save({k: v.cpu() for k, v in got.items()}, path)
else:
    ref = torch.load(path)
    for k in ids:
        a, b = got[k].float().cpu().flatten(), ref[k].float().flatten()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        line = "%-45s cos %.6f  max|d| %.3e  max|ref| %.3e" % (k, cos, (a - b).abs().max().item(), b.abs().max().item())
        if k.endswith("-map"):     # every (image, head) block on its own: a wrong batch / head stride would show here
            ga, rb = got[k].double().cpu(), ref[k].double()
            per = torch.nn.functional.cosine_similarity(ga.flatten(2), rb.flatten(2), dim=2)
            line += "  min cos over (b, head) %.6f" % per.min().item()
        print(line)
