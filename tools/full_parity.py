"""Full-size parity: SDXL 1024^2 (or --img), batch 1, all 472 non-map activations, CUDA path vs the CPU oracle on the
same synthetic weights / inputs / injected noise. Writes a per-map report (cosine, relative L2, max-relative).
    python tools/full_parity.py --out gpurun_out/full_parity.json [--img 1024] [--version xl]
"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import O, build_oracle, compare_maps, make_inputs
from generic_diffusion_feature_b200.components import models
from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor

ap = argparse.ArgumentParser()
ap.add_argument("--img", type=int, default=1024)
ap.add_argument("--version", default="xl")
ap.add_argument("--out", default="gpurun_out/full_parity.json")
a = ap.parse_args()
ver = a.version
ucfg, vcfg = models.UNET_CONFIGS[ver], models.VAE_CONFIGS[ver]
torch.set_num_threads(os.cpu_count())
t0 = time.time()
sd = models.synthetic_state_dict(ver, "cuda:0")
print("weights %.1fs" % (time.time() - t0), flush=True)
pooled_dim = 1280 if ucfg["add_time_dim"] else None
image, ctx, pooled, ev, eq = make_inputs(1, a.img, ucfg["ctx_dim"], pooled_dim)
ids = _unet_feature_ids(ucfg)
layer = {i: True for i in ids}
pipe = models.get_diffusion_model(ver, "float16", device="cuda:0", state_dict=sd)
fe = FeatureExtractor(layer, ver, "cuda:0", img_size=a.img, external_model=pipe)
got = fe.extract((ctx, ctx, pooled, pooled), 1, image.cuda(), image_type="tensors", t=50, noise=(ev, eq))
torch.cuda.synchronize()
got = {k: v.float().cpu() for k, v in got.items()}
print("cuda done %.1fs" % (time.time() - t0), flush=True)
sd_cpu = {k: v.cpu() for k, v in sd.items()}
del sd
unet, vae = build_oracle(ucfg, vcfg, sd_cpu)
store = O.FeatureStore(layer)
O.attach_gatherers(unet, store)
t1 = time.time()
want, _, _ = O.extract(ver, unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=a.img)
print("oracle %.1fs on %d cores" % (time.time() - t1, os.cpu_count()), flush=True)
rows = compare_maps(got, want)
worst = sorted(rows, key=lambda r: r[1])[:10]
rep = {"version": ver, "img": a.img, "maps": len(rows), "min_cos": min(r[1] for r in rows),
       "median_cos": sorted(r[1] for r in rows)[len(rows) // 2], "max_rel_l2": max(r[2] for r in rows),
       "max_maxrel": max(r[3] for r in rows), "n_below_0.999": sum(r[1] < 0.999 for r in rows),
       "worst10": worst, "oracle_seconds": time.time() - t1, "cores": os.cpu_count(),
       "rows": rows}
json.dump(rep, open(a.out, "w"))
print(json.dumps({k: v for k, v in rep.items() if k != "rows"}, indent=1))
