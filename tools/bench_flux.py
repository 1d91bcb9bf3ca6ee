"""Full-size FLUX.1-dev forward with capture on one B200 (synthetic weights by name, 12 B parameters): time of
`FeatureExtractor.extract` at 1024x1024, batch 1, t = 50, all 323 maps (19 x 7 + 38 x 5) captured.
    python tools/bench_flux.py [--steps K] > profiles/rNN_flux_full.json
Not the headline benchmark (BASELINE.json names SDXL); a measured data point for SURVEY.md row a17."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.components.feature_extractor import _flux_feature_ids
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor
from generic_diffusion_feature_b200 import _lib
import ctypes

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--img", type=int, default=1024)
ap.add_argument("--layers", type=int, default=19)
ap.add_argument("--single", type=int, default=38)
a = ap.parse_args()
dev = "cuda:0"
fcfg = dict(models.FLUX_CONFIGS["flux"], layers=a.layers, single_layers=a.single)
t0 = time.time()
pipe = models.get_diffusion_model("flux", "float16", device=dev, weight_device=dev, flux_cfg=fcfg,
                                  vae_cfg=models.VAE_CONFIGS["flux"], synthetic=True)
t_load = time.time() - t0
ids = _flux_feature_ids(fcfg)
fe = FeatureExtractor({i: True for i in ids}, "flux", dev, img_size=a.img, external_model=pipe)
g = torch.Generator(device=dev).manual_seed(1234)
image = torch.rand(1, 3, a.img, a.img, generator=g, device=dev) * 2 - 1
L = a.img // 8
noise = (torch.randn(1, 16, L, L, generator=g, device=dev), torch.randn(1, 16, L, L, generator=g, device=dev))
prompts = tuple(p.to(dev) for p in fe.encode_prompt(""))
for _ in range(2):
    feats = fe.extract(prompts, 1, image, image_type="tensors", t=50, noise=noise)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    feats = fe.extract(prompts, 1, image, image_type="tensors", t=50, noise=noise)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
finite = all(bool(torch.isfinite(v).all()) for v in list(feats.values())[::40])
lib = pipe.lib
_lib.check(lib.gdf_profile(pipe.handle, 1))
fe.extract(prompts, 1, image, image_type="tensors", t=50, noise=noise)
torch.cuda.synchronize()
msv = (ctypes.c_float * 5)(); flv = (ctypes.c_double * 5)(); lnv = (ctypes.c_int * 5)()
_lib.check(lib.gdf_profile_read(pipe.handle, msv, flv, lnv))
_lib.check(lib.gdf_profile(pipe.handle, 0))
kinds = ["tcgen05_gemm_conv", "attention", "groupnorm", "layernorm", "other"]
n_params = sum(int(torch.tensor(s).prod()) for _, s in models.flux_param_specs(fcfg))
print(json.dumps({"workload": "FLUX.1-dev-sized MMDiT (%d double + %d single blocks) %dx%d, batch 1, %d maps captured"
                  % (a.layers, a.single, a.img, a.img, len(ids)), "ms_per_image": ms, "images_per_s": 1e3 / ms,
                  "transformer_params": n_params, "arena_gb": fe._plan.arena_bytes / 1e9, "launches": fe._plan.launches,
                  "weight_load_s": t_load, "finite": finite,
                  "per_kind_ms": {k: msv[i] for i, k in enumerate(kinds)},
                  "per_kind_tflops": {k: (flv[i] / (msv[i] * 1e-3) / 1e12 if msv[i] > 0 else 0) for i, k in enumerate(kinds)}}))
