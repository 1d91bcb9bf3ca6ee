#!/usr/bin/env python
"""Benchmark of the extraction hot path (BASELINE.json): images/s (pairs/s) of the five named configurations and
the achieved bandwidth of the HBM-bound kernels.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--batch B]

  --config sdxl_1024 (default, BASELINE configs[1], the headline): SDXL UNet 1024x1024, all 472 non-`map` activations
           of config_xl_full captured at native resolution into the fp16 arena, bf16, batch 8 / GPU
  --config sd15_512     configs[0]: SD-1.5 512x512, batch 1, t = 50, the reference's config_15_legacy.json (4 maps)
  --config sd21_768_mt  configs[2]: SD-2.1 768x768, t in {50, 150, 250}, all 165 maps per timestep, every map resized
                        bilinearly to 96x96 and concatenated (aggregation_network.py:62-66): img/s + stack GB/s
  --config pixart_1024  configs[3]: PixArt-Sigma 1024x1024, 28 x {self-q,k,v, cross-q, ffn-inner, out} = 168 maps
  --config corr_sdxl    configs[4]: SDXL `practical` features (3840 channels) of synthetic image pairs -> stacks at
                        128x128 -> cosine-similarity arg-max of 4096 query points at load size 512: pairs/s
  --config hbm_kernels  every HBM-bound kernel of the path timed alone on tensors larger than L2: GB/s vs the measured
                        copy bandwidth (MEASURED_PEAKS.json)

One "step" = one pass of the hot path over one batch of synthetic inputs (weights random-init by parameter name,
images / conditioning / noise synthetic: there is no network for checkpoints or datasets).

Rank 0 prints ONE JSON line (contract in the task statement): `value` = whole-job throughput with inputs resident in
HBM (CUDA events, max over ranks), `e2e` = the same through the public API with pinned-host inputs (H2D inside the timed
region) and a D2H read of the step's result, `roofline` for the dominant kernel from a CUDA-event profiling pass,
`cpu_baseline` = the CPU oracle port timed on the host cores on a bounded sample (N = 1 only).
`--impl reference` times the CPU oracle port (the reference's diffusers path cannot be imported: diffusers is neither
installed nor installable offline - probed on the GPU box too) on the same config.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KINDS = ["tcgen05_gemm_conv", "attention", "groupnorm", "layernorm", "other"]

# per configuration: metric name, unit, default batch per GPU, algorithmic FLOPs per unit, workload text
CONFIGS = {
    "sdxl_1024": dict(metric="sdxl_1024_full_activation_extraction_images_per_s", unit="images/s", batch=8,
                      flop_per_unit=11.640e12, version="xl", img=1024,
                      workload="SDXL UNet 1024x1024 full activation set (472 maps, fp16 arena), VAE encode + q_sample "
                               "t=50"),
    "sd15_512": dict(metric="sd15_512_legacy_extraction_images_per_s", unit="images/s", batch=1,
                     flop_per_unit=1.920e12, version="1-5", img=512,
                     workload="SD-1.5 UNet 512x512, config_15_legacy (4 up-block maps, 3520 channels), VAE encode + "
                              "q_sample t=50"),
    "sd21_768_mt": dict(metric="sd21_768_multi_timestep_stack_images_per_s", unit="images/s", batch=8,
                        flop_per_unit=3 * (2.149e12 + 2.609e12), version="2-1", img=768,
                        workload="SD-2.1 UNet 768x768, t in {50,150,250} (VAE re-encoded per timestep like the "
                                 "reference), all 165 maps per timestep, every map bilinear -> 96x96 + channel concat "
                                 "(170888 channels per timestep)"),
    "pixart_1024": dict(metric="pixart_sigma_1024_block_capture_images_per_s", unit="images/s", batch=8,
                        flop_per_unit=11.513e12, version="pixart-sigma", img=1024,
                        workload="PixArt-Sigma DiT 1024x1024, 28 blocks x {self-q,k,v, cross-q, ffn-inner, out} = 168 "
                                 "maps, 300 caption tokens, VAE encode + q_sample t=50"),
    "corr_sdxl": dict(metric="sdxl_pair_correspondence_pairs_per_s", unit="pairs/s", batch=8,
                      flop_per_unit=2 * 11.640e12 + 0.515e12, version="xl", img=1024,
                      workload="SDXL practical features (3840 ch) of image pairs -> stacks 128x128 -> cosine-similarity "
                               "arg-max, 4096 queries, load size 512 (exact low-resolution form)"),
    "hbm_kernels": dict(metric="hbm_kernel_bandwidth_resize_concat_gbs", unit="GB/s", batch=8, flop_per_unit=0.0,
                        version=None, img=0,
                        workload="HBM-bound kernels of the path timed alone (resize+concat, LayerNorm, GroupNorm, "
                                 "avg-pool, q_sample, nearest x2, Flux qk-norm+RoPE) on tensors larger than L2"),
}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1371.9), d.get("bf16_tflops", 1668.4), d.get("hbm_gbs", 6540.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def read_traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu launch list of this same command
    (profiles/r02_traffic.json: {config: {kernel, dram_bytes_per_launch, source}}); None if absent."""
    for fn in ("r02_traffic.json",):
        p = os.path.join(ROOT, "profiles", fn)
        if os.path.exists(p):
            try:
                return json.load(open(p)).get(name)
            except Exception:  # noqa: BLE001
                return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner through
    NCCL_DEBUG_FILE = stdout when the box sets NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the
    whole run and the JSON line is written to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------ layer selections
def _ref_ids(name):
    """The reference's own layer JSONs (feature/configs/*.json), committed as tests/golden/reference_ids.json."""
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_ids.json")))[name]


def layer_for(config):
    from generic_diffusion_feature_b200.components import feature_extractor as fx
    from generic_diffusion_feature_b200.components.models import DIT_CONFIGS, UNET_CONFIGS
    if config == "sdxl_1024":
        return {i: True for i in fx._unet_feature_ids(UNET_CONFIGS["xl"])}   # == non-map ids of config_xl_full.json
    if config == "sd15_512":
        return _ref_ids("config_15_legacy.json")
    if config == "sd21_768_mt":
        return {i: True for i in fx._unet_feature_ids(UNET_CONFIGS["2-1"])}
    if config == "pixart_1024":
        return {i: True for i in fx._dit_feature_ids(DIT_CONFIGS["pixart-sigma"])}
    if config == "corr_sdxl":
        return _ref_ids("config_xl_practical.json")
    raise ValueError(config)


def full_xl_layer():
    return layer_for("sdxl_1024")


QUERY_SEED = 1239


def query_points(n=4096, load=512):
    import numpy as np
    return np.random.RandomState(QUERY_SEED).uniform(0, load - 1, size=(n, 2))


# ------------------------------------------------------------------------------------------ CPU oracle (reference arm)
def cpu_oracle_rate(config, sd, steps, warmup):
    """CPU port of the reference path (oracle) on the host cores for one bounded sample of `config`:
    returns (units/s, cores, seconds per sample, sample description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import O, build_oracle, build_oracle_dit, make_dit_inputs, make_inputs
    from generic_diffusion_feature_b200.components.models import DIT_CONFIGS, UNET_CONFIGS, VAE_CONFIGS
    cfg = CONFIGS[config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    layer = layer_for(config)
    img = cfg["img"]
    if config == "pixart_1024":
        model, vae = build_oracle_dit(DIT_CONFIGS["pixart-sigma"], VAE_CONFIGS["pixart-sigma"], sd)
        store = O.FeatureStore(layer)
        O.attach_gatherers_dit(model, store)
        image, ctx, mask, ev, eq = make_dit_inputs(1, img, 4096, ctx_len=300, masked_tail=0)
        run = lambda: O.extract_dit("pixart-sigma", model, vae, store, image, ctx, mask, ev, eq, t=50)
        units, what = 1, "1 image 1024x1024, 168 maps"
    else:
        v = cfg["version"]
        unet, vae = build_oracle(UNET_CONFIGS[v], VAE_CONFIGS[v], sd)
        store = O.FeatureStore(layer)
        O.attach_gatherers(unet, store)
        ctx_dim = UNET_CONFIGS[v]["ctx_dim"]
        pooled_dim = 1280 if v == "xl" else None
        if config == "corr_sdxl":
            image, ctx, pooled, ev, eq = make_inputs(2, img, ctx_dim, pooled_dim)
            pts = query_points()

            def run():
                feats, _, _ = O.extract(v, unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=img)
                stack = O.resize_concat(list(feats.values()), (128, 128))
                # reference formulation (correspondence_utils.py:113-138): both stacks upsampled to 512x512, full
                # similarity matrix; 512 queries at a time keep the fp32 matrix at 0.5 GB
                out = []
                for c0 in range(0, len(pts), 512):
                    p2, _ = O.find_nn_source_correspondences(stack[0:1], stack[1:2], pts[c0:c0 + 512], (512, 512))
                    out.append(p2)
                return torch.cat(out)
            units, what = 1, "1 pair (2 images 1024x1024, 3840-channel stacks, 4096 queries, reference formulation)"
        elif config == "sd21_768_mt":
            image, ctx, pooled, ev, eq = make_inputs(1, img, ctx_dim, pooled_dim)

            def run():
                for t in (50, 150, 250):
                    feats, _, _ = O.extract(v, unet, vae, store, image, ctx, pooled, ev, eq, t=t, img_size=img)
                    O.resize_concat(list(feats.values()), (96, 96))
            units, what = 1, "1 image 768x768 x 3 timesteps, 165 maps each, resize + concat to 96x96"
        else:
            image, ctx, pooled, ev, eq = make_inputs(1, img, ctx_dim, pooled_dim)
            run = lambda: O.extract(v, unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=img)
            units, what = 1, "1 image %dx%d, %d maps" % (img, img, len(layer))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return units / sec, cores, sec, what


def _state_dict_cpu(config, dev="cpu"):
    from generic_diffusion_feature_b200.components.models import synthetic_state_dict
    sd = synthetic_state_dict(CONFIGS[config]["version"], dev)
    return {k: v.cpu() for k, v in sd.items()}


def config_block(config, batch, world, extra=None):
    cfg = CONFIGS[config]
    c = {"workload": cfg["workload"], "name": config, "batch_per_gpu": batch, "global_batch": batch * world,
         "parallelism": "dp%d (replicated weights, inputs sharded, no data-path collective)" % world}
    if extra:
        c.update(extra)
    return c


def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    if args.config == "hbm_kernels":
        emit({"impl": "reference", "metric": cfg["metric"], "unavailable":
              "kernel micro-benchmark of this repo's own kernels: no reference counterpart"})
        return
    sd = _state_dict_cpu(args.config)
    rate, cores, sec, what = cpu_oracle_rate(args.config, sd, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": rate, "unit": cfg["unit"], "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_block(args.config, 1, 1, {"parallelism": "host threads (%d)" % cores,
                                                   "sample_per_step": what}),
        "cpu_baseline": {"value": rate, "unit": cfg["unit"], "cores": cores, "kind": "port",
                         "sample": what + " per step through the CPU oracle (torch fp32 restatement of the reference "
                                          "path; diffusers itself is not installable offline)"},
        "e2e": {"value": rate, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ workloads (GPU)
class Workload:
    """One configuration on one rank: device-resident step, end-to-end step (pinned-host inputs, D2H of the result)."""

    def __init__(self, config, dev, rank, batch):
        from generic_diffusion_feature_b200.components.models import DIT_CONFIGS, UNET_CONFIGS, get_diffusion_model
        from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor
        self.config, self.dev, self.B = config, dev, batch
        cfg = CONFIGS[config]
        v, img = cfg["version"], cfg["img"]
        self.img = img
        self.pipe = get_diffusion_model(v, "float16", device=dev, weight_device=dev, synthetic=True)
        self.layer = layer_for(config)
        self.fe = FeatureExtractor(self.layer, v, dev, img_size=img, external_model=self.pipe)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        B, L = batch, img // 8
        self.images = torch.rand(B, 3, img, img, generator=g, device=dev) * 2 - 1
        if config == "corr_sdxl":   # second image of every pair = shifted, dimmed copy of the first (meaningful matches)
            self.images[1::2] = torch.roll(self.images[0::2], shifts=(24, -16), dims=(2, 3)) * 0.9 + 0.1 * self.images[1::2]
        self.noise = (torch.randn(B, 4, L, L, generator=g, device=dev), torch.randn(B, 4, L, L, generator=g, device=dev))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")     # stand-in prompt embeddings are what a synthetic benchmark wants
            self.prompts = tuple(None if p is None else p.to(dev) for p in self.fe.encode_prompt(""))
        self.images_host = self.images.cpu().pin_memory()
        self.h2d = self.images_host.numel() * 4
        self.units_per_step = B if cfg["unit"] == "images/s" else B // 2
        self.stack_ms = []          # CUDA-event times of the resize+concat launches (sd21_768_mt, corr_sdxl)
        self.stack_bytes = 0
        self.corr_ms = []
        self._ev = []
        self.timesteps = (50, 150, 250) if config == "sd21_768_mt" else (50,)
        if config == "corr_sdxl":
            import numpy as np
            from generic_diffusion_feature_b200 import correspondence as C
            self.pts = query_points()
            self.C = C
        self.d2h = 0

    # -- pieces
    def _extract(self, images, t):
        return self.fe.extract(self.prompts, self.B, images, image_type="tensors", t=t, noise=self.noise)

    def _timed(self, bucket, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self._ev.append((bucket, e0, e1))
        return out

    def collect_events(self):
        for bucket, e0, e1 in self._ev:
            bucket.append(e0.elapsed_time(e1))
        self._ev = []

    def _stack(self, feats, hw, layout="nhwc", sumsq=False):
        from generic_diffusion_feature_b200 import ops
        maps = []
        for f in feats.values():
            B, C, h, w = f.shape
            maps.append(f.permute(0, 2, 3, 1).reshape(B, h * w, C))     # the arena's token-major storage, no copy
        # maps whose channel count is not a multiple of 8 (the 4-channel unet-in / unet-out) go last, so that every other
        # map keeps 16-byte aligned channel offsets in the stack (a consumer reads the stack through its own id -> offset
        # table either way; the reference concatenates in dict order, aggregation_network.py:62-66)
        maps.sort(key=lambda m: m.shape[2] % 8 != 0)
        self.stack_bytes = sum(m.numel() * 2 for m in maps) + self.B * hw * hw * sum(m.shape[2] for m in maps) * 2
        return self._timed(self.stack_ms, lambda: ops.resize_concat(maps, (hw, hw), nhwc=(layout == "nhwc"),
                                                                    nchw=(layout == "nchw"), with_sumsq=sumsq))

    def _step(self, images):
        c = self.config
        if c == "sd21_768_mt":
            out = None
            for t in self.timesteps:
                feats = self._extract(images, t)
                out = self._stack(feats, 96)["nhwc"]     # one stack buffer per timestep (25 GB at B = 8), reused
            return out
        if c == "corr_sdxl":
            feats = self._extract(images, 50)
            stack = self._stack(feats, 128)["nhwc"]       # [B, 16384, 3840] fp16
            res = []
            for pr in range(self.B // 2):
                res.append(self._timed(self.corr_ms, lambda pr=pr: self.C.find_nn_source_correspondences(
                    stack[2 * pr:2 * pr + 1], stack[2 * pr + 1:2 * pr + 2], self.pts, None, (512, 512))[1]))
            return torch.stack(res)
        return self._extract(images, 50)

    def step_device(self):
        return self._step(self.images)

    def step_e2e(self):
        # Every step issues ONE pinned host -> device copy of a full input batch and one device -> host read of its result,
        # both inside the timed region. The copy runs on the extractor's side stream (FeatureExtractor.stage_images, the
        # public prefetch hook): the batch a step consumes was put in flight by the previous step, the one it issues
        # overlaps its own forward - the way a caller looping over a dataset uses the API (and the CLI does).
        if getattr(self, "_staged", None) is None:
            self._staged = self.fe.stage_images(self.images_host)
        images, self._staged = self._staged, None
        out = self._step(images)
        self._staged = self.fe.stage_images(self.images_host)
        if self.config == "sd21_768_mt":
            r = out[:, :, :4].contiguous().cpu()          # D2H read of the step's result (syncs the step)
        elif self.config == "corr_sdxl":
            r = out.cpu()
        else:
            key = "unet-out" if "unet-out" in out else list(out.keys())[-1]
            r = out[key].contiguous().cpu() if self.config != "pixart_1024" else out[key][:, :8].contiguous().cpu()
        self.d2h = r.numel() * r.element_size()
        return r

    def launches_per_step(self):
        n = (self.fe._plan.launches + 3) * len(self.timesteps)
        if self.config == "sd21_768_mt":
            n += 2 * len(self.timesteps)
        if self.config == "corr_sdxl":
            n += 2 + 8 * (self.B // 2)
        return n


def profile_pass(wl, lib, handle):
    """CUDA-event profiling pass of the library (one event pair per kernel launch on the launching stream)."""
    from generic_diffusion_feature_b200 import _lib
    _lib.check(lib.gdf_profile(handle, 1))
    reps = 2
    for _ in range(reps):
        wl.step_device()
    torch.cuda.synchronize()
    msv, flv, lnv = (ctypes.c_float * 5)(), (ctypes.c_double * 5)(), (ctypes.c_int * 5)()
    _lib.check(lib.gdf_profile_read(handle, msv, flv, lnv))
    return reps, msv, flv, lnv


def run_ours(args, rank, world, local):
    import torch.distributed as dist
    from generic_diffusion_feature_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the extraction path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line of the contract
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["batch"]
    if args.config == "corr_sdxl" and B % 2:
        raise SystemExit("corr_sdxl works on image pairs: --batch must be even")
    if args.config == "hbm_kernels":
        return run_hbm_kernels(args, rank, world, dev)
    wl = Workload(args.config, dev, rank, B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    W = max(args.warmup, 3)
    for _ in range(W):
        wl.step_device()
    torch.cuda.synchronize()
    wl._ev, wl.stack_ms, wl.corr_ms = [], [], []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.cuda_profiler:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: capture the timed steps only
    ms = timed(wl.step_device, args.steps)
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    wl.collect_events()
    stack_ms, corr_ms = list(wl.stack_ms), list(wl.corr_ms)
    value = world * wl.units_per_step * args.steps / (ms / 1e3)

    for _ in range(2):
        wl.step_e2e()
    ms_e2e = timed(wl.step_e2e, args.steps)
    wl._ev = []
    e2e_value = world * wl.units_per_step * args.steps / (ms_e2e / 1e3)

    # gathered results on one rank (north_star: NCCL over NVLink only where a consumer wants them on one rank):
    # correspondence indices of every pair (32 KB / pair) go to rank 0 through point-to-point sends
    gather = None
    if args.config == "corr_sdxl" and world > 1:
        from generic_diffusion_feature_b200 import parallel
        res = wl.step_device().contiguous()
        counts = [res.shape[0]] * world
        for _ in range(2):
            parallel.gather_to_rank0(res, counts)
        gms = timed(lambda: parallel.gather_to_rank0(res, counts), 10) / 10
        gb = parallel.gather_bytes(counts, res[0].numel() * res.element_size())
        gather = {"what": "arg-max indices of all pairs to rank 0 (dist.send / irecv, NCCL)", "bytes": gb,
                  "ms": gms, "gbs": gb / (gms * 1e-3) / 1e9}
    if args.config == "sdxl_1024" and world > 1 and args.gather_stacks:
        # optional: the consumer-side gather of north_star at its largest sensible size - one 128x128x3840 fp16 stack
        # per image (126 MB) from every rank to rank 0
        from generic_diffusion_feature_b200 import parallel
        st = torch.empty(B, 128 * 128, 3840, dtype=torch.float16, device=dev).normal_()
        counts = [B] * world
        for _ in range(2):
            parallel.gather_to_rank0(st, counts)
        gms = timed(lambda: parallel.gather_to_rank0(st, counts), 5) / 5
        gb = parallel.gather_bytes(counts, st[0].numel() * 2)
        gather = {"what": "128x128x3840 fp16 stacks of every image to rank 0 (dist.send / irecv, NCCL)", "bytes": gb,
                  "ms": gms, "gbs": gb / (gms * 1e-3) / 1e9}

    # ---- roofline of the dominant kernel: CUDA-event profiling pass on the launching stream
    lib, roof = wl.pipe.lib, None
    sustained, burst, hbm, how = read_peaks()
    if rank == 0:
        reps, msv, flv, lnv = profile_pass(wl, lib, wl.pipe.handle)
        if args.profile_csv:
            _lib.check(lib.gdf_profile_dump(wl.pipe.handle, args.profile_csv.encode()))
        _lib.check(lib.gdf_profile(wl.pipe.handle, 0))
        wl._ev = []
        n_pass = reps * len(wl.timesteps)
        gemm_tflops = flv[0] / (msv[0] * 1e-3) / 1e12 if msv[0] > 0 else 0.0
        tot_ms = sum(msv)
        traffic = read_traffic(args.config)
        per_unit_flops = cfg["flop_per_unit"]
        whole = per_unit_flops * value / world / 1e12
        tensor_roof = {"bound": "tensor", "achieved": gemm_tflops, "peak": sustained, "unit": "TFLOP/s",
                       "frac": gemm_tflops / sustained,
                       "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                       "traffic_source": traffic["source"] if traffic else None,
                       "algorithmic_flop_per_launch": flv[0] / max(lnv[0], 1),
                       "peak_source": how + " (sustained bf16, cuBLAS 8192^3 back to back)",
                       "kernel": "gemm_tcgen05_kernel",
                       "avg_launch_us": 1e3 * msv[0] / max(lnv[0], 1),
                       "share_of_step": msv[0] / tot_ms if tot_ms > 0 else None,
                       "per_kind_ms_per_step": {k: msv[i] / reps for i, k in enumerate(KINDS)},
                       "per_kind_launches_per_step": {k: lnv[i] // reps for i, k in enumerate(KINDS)},
                       "attention_tflops": (flv[1] / (msv[1] * 1e-3) / 1e12) if msv[1] > 0 else None,
                       "whole_path_tflops": whole, "whole_path_frac": whole / sustained}
        roof = tensor_roof
        if args.config == "sd21_768_mt" and stack_ms:
            # the bound BASELINE names for this config's second half: resize+concat is HBM bound
            avg = sum(stack_ms) / len(stack_ms)
            gbs = wl.stack_bytes / (avg * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                    "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                    "traffic_source": traffic["source"] if traffic else None,
                    "kernel": "resize_concat_tile_kernel (bilinear resize of 165 maps to 96x96 + channel concat, 3 launches of <= 64 maps)",
                    "algorithmic_bytes_per_launch": wl.stack_bytes, "avg_launch_us": avg * 1e3,
                    "launches_timed": len(stack_ms), "peak_source": how + " (copy bandwidth)",
                    "share_of_step": (avg * len(wl.timesteps)) / (ms / args.steps),
                    "tensor": tensor_roof}
        if args.config == "corr_sdxl":
            roof = dict(tensor_roof)
            if stack_ms:
                avg = sum(stack_ms) / len(stack_ms)
                roof["stack"] = {"bound": "hbm", "achieved": wl.stack_bytes / (avg * 1e-3) / 1e9, "peak": hbm,
                                 "unit": "GB/s", "algorithmic_bytes_per_launch": wl.stack_bytes, "avg_launch_us": avg * 1e3}
                roof["stack"]["frac"] = roof["stack"]["achieved"] / hbm
            if corr_ms:
                avg = sum(corr_ms) / len(corr_ms)
                roof["correspond"] = {"ms_per_pair": avg, "executed_tflop_per_pair": 0.515,
                                      "tflops": 0.515 / (avg * 1e-3), "reference_formulation_tflop_per_pair": 8.246}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            rate, cores, sec, what = cpu_oracle_rate(args.config, _state_dict_cpu(args.config, dev), 1, 0)
            cpu_base = {"value": rate, "unit": cfg["unit"], "cores": cores, "kind": "port",
                        "sample": "%s, CPU oracle (torch fp32), %.1f s" % (what, sec)}
        except Exception as ex:  # noqa: BLE001
            cpu_base = {"value": None, "unit": cfg["unit"], "cores": os.cpu_count(), "kind": "port",
                        "sample": "failed: %r" % (ex,)}

    if rank == 0:
        plan = wl.fe._plan
        line = {
            "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_block(args.config, B, world, {
                "l2": "working set per step (weights + %.1f GB arena) >> 126 MB L2, no flush needed"
                      % (plan.arena_bytes / 1e9),
                "arena_gb": plan.arena_bytes / 1e9, "workspace_gb": plan.workspace_bytes / 1e9,
                "units_per_step_per_gpu": wl.units_per_step}),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": cfg["unit"], "h2d_bytes_per_step": wl.h2d,
                    "d2h_bytes_per_step": wl.d2h, "ms_per_step": ms_e2e / args.steps,
                    "overlap": "every step issues one pinned H2D of a full batch (FeatureExtractor.stage_images, side "
                               "stream) that overlaps its own forward and feeds the next step, and one D2H read of its "
                               "result; both inside the timed region"},
            "gpu_launches": wl.launches_per_step() * args.steps,
            "roofline": roof, "cpu_baseline": cpu_base,
        }
        if gather:
            line["gather"] = gather
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ HBM-bound kernels
def run_hbm_kernels(args, rank, world, dev):
    """Every HBM-bound kernel of the path timed alone with CUDA events: algorithmic bytes (each input read once, each
    output written once) / time vs the measured copy bandwidth. Shapes are the ones of the SDXL-1024 / SD-2.1-768 /
    Flux steps at batch 8; every working set is larger than the 126 MB L2 unless noted."""
    from generic_diffusion_feature_b200 import correspondence as C
    from generic_diffusion_feature_b200 import ops
    if rank != 0:
        return
    sustained, burst, hbm, how = read_peaks()
    g = torch.Generator(device=dev).manual_seed(7)
    rnd = lambda *s: torch.randn(*s, generator=g, device=dev)
    rows = []

    def bench(name, bytes_, fn, note="", iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        rows.append({"kernel": name, "us": us, "algorithmic_bytes": bytes_, "gbs": bytes_ / us * 1e-3,
                     "frac": bytes_ / us * 1e-3 / hbm, "note": note})

    B = args.batch or 8
    # ceilings measured live on this GPU: fill (write only) and copy (read + write) of 1 GB. HBM3e write-only traffic tops
    # out well below the copy figure (3.9 TB/s against 6.6 TB/s on the B200s of this pool), and the stack kernels write
    # 4x what they read, so their fraction of the WRITE ceiling (`write_frac`) is the one that says how far they are
    # from the hardware; `frac` stays algorithmic bytes / time against the copy peak of MEASURED_PEAKS.json.
    ceil_buf = torch.empty(512 * 1024 * 1024, dtype=torch.float16, device=dev)
    ceil_dst = torch.empty_like(ceil_buf)
    bench("ceiling: fill 1 GB (write only, torch zero_)", ceil_buf.numel() * 2, lambda: ceil_buf.zero_(), "reference point")
    bench("ceiling: copy 1 GB (read + write, torch copy_)", 2 * ceil_buf.numel() * 2, lambda: ceil_dst.copy_(ceil_buf),
          "reference point")
    write_ceiling = rows[0]["gbs"]
    del ceil_buf, ceil_dst
    # resize + concat: SDXL practical maps (B x 3840 channels at 32 / 64 / 128) -> 128x128 stack, NHWC (+ per-pixel norms)
    maps = [rnd(B, 32 * 32, 1280).half(), rnd(B, 32 * 32, 1280).half(), rnd(B, 64 * 64, 640).half(),
            rnd(B, 128 * 128, 320).half(), rnd(B, 128 * 128, 320).half()]
    ctot = sum(m.shape[2] for m in maps)
    by = sum(m.numel() * 2 for m in maps) + B * 128 * 128 * ctot * 2
    wr = B * 128 * 128 * ctot * 2
    bench("resize_concat_tile_kernel (3840 ch -> 128x128, + sumsq)", by,
          lambda: ops.resize_concat(maps, (128, 128), nhwc=True, with_sumsq=True), "SDXL practical stack, B=%d" % B)
    rows[-1]["write_frac"] = wr / rows[-1]["us"] * 1e-3 / write_ceiling
    bench("resize_nchw_kernel (3840 ch -> 128x128, reference layout)", by,
          lambda: ops.resize_concat(maps, (128, 128), nhwc=False, nchw=True), "smem transpose to (B,C,H,W)")
    rows[-1]["write_frac"] = wr / rows[-1]["us"] * 1e-3 / write_ceiling
    # SD-2.1 768 multi-timestep stack: many maps -> 96x96
    m21 = [rnd(B, 96 * 96, 320).half() for _ in range(6)] + [rnd(B, 48 * 48, 640).half() for _ in range(6)] + \
          [rnd(B, 24 * 24, 1280).half() for _ in range(8)] + [rnd(B, 12 * 12, 1280).half() for _ in range(6)]
    c21 = sum(m.shape[2] for m in m21)
    by21 = sum(m.numel() * 2 for m in m21) + B * 96 * 96 * c21 * 2
    bench("resize_concat_tile_kernel (26 maps, %d ch -> 96x96)" % c21, by21, lambda: ops.resize_concat(m21, (96, 96), nhwc=True),
          "SD-2.1 768 `-out` maps")
    rows[-1]["write_frac"] = (B * 96 * 96 * c21 * 2) / rows[-1]["us"] * 1e-3 / write_ceiling
    # LayerNorm at the two SDXL transformer shapes (21 / 42 MB tensors: L2 resident in the step, as in the model)
    for M, Cc in ((8192, 1280), (32768, 640), (262144, 1152)):
        x = rnd(M, Cc).bfloat16()
        gm, bt = rnd(Cc), rnd(Cc)
        bench("layernorm_rows_kernel M=%d C=%d" % (M, Cc), 2 * M * Cc * 2, lambda x=x, gm=gm, bt=bt: ops.layernorm(x, gm, bt, 1e-5),
              "L2-resident" if M * Cc * 4 < 120e6 else "")
    # GroupNorm + SiLU at the VAE shapes (2 reads + 1 write of the tensor)
    for HW, Cc in ((1024 * 1024, 128), (512 * 512, 256), (128 * 128, 512), (128 * 128, 320), (32 * 32, 1280)):
        x = rnd(B, HW, Cc).bfloat16()
        gm, bt = rnd(Cc), rnd(Cc)
        bench("groupnorm (stats + apply) HW=%d C=%d" % (HW, Cc), 3 * x.numel() * 2,
              lambda x=x, gm=gm, bt=bt: ops.groupnorm(x, gm, bt, 32, 1e-6, True), "B=%d" % B)
        del x
    # feature_resize average pooling of a captured map
    x = rnd(B, 128, 128, 640).half()
    y = torch.empty(B, 64, 64, 640, dtype=torch.float16, device=dev)
    from generic_diffusion_feature_b200 import _lib
    lib = _lib.load()
    bench("adaptive_avgpool_nhwc_kernel 128x128x640 -> 64x64", x.numel() * 2 + y.numel() * 2,
          lambda: _lib.check(lib.gdf_op_avgpool_nhwc(_lib.ptr(x), _lib.ptr(y), B, 128, 128, 640, 64, 64,
                                                     _lib.stream_ptr())))
    # nearest x2 upsampling (UNet up path)
    x = rnd(B, 64, 64, 640).bfloat16()
    bench("upsample_nearest2x_kernel 64x64x640", x.numel() * 2 * 5, lambda: ops.upsample_nearest2x(x))
    # correspondence on 128x128x3840 stacks (tensor + HBM mix; reported as time per pair)
    s = rnd(2, 128 * 128, 3840).half()
    pts = query_points()
    bench("gdf_correspond (4096 queries, 128x128x3840 stacks, load 512)", 2 * s[0].numel() * 2,
          lambda: C.find_nn_source_correspondences(s[0:1], s[1:2], pts, None, (512, 512)),
          "similarity GEMM 0.515 TFLOP + norm map + interpolated arg-max; bytes = the two stacks read once", iters=10)
    top = rows[2]
    line = {"metric": CONFIGS["hbm_kernels"]["metric"], "value": top["gbs"], "unit": "GB/s", "n_gpus": 1,
            "steps": 20, "warmup": 3, "ms_per_step": top["us"] * 1e-3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": config_block("hbm_kernels", B, 1, {"l2": "inputs larger than the 126 MB L2 unless noted"}),
            "roofline": {"bound": "hbm", "achieved": top["gbs"], "peak": hbm, "unit": "GB/s", "frac": top["frac"],
                         "traffic": None, "kernel": top["kernel"], "peak_source": how + " (copy bandwidth)",
                         "write_only_ceiling_gbs": write_ceiling, "write_frac": top.get("write_frac")},
            "kernels": rows, "gpu_launches": len(rows) * 23, "cpu_baseline": None,
            "e2e": {"value": top["gbs"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "kernel micro-benchmark: operands resident in HBM by construction"}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="sdxl_1024", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the configuration's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-stacks", action="store_true", help="sdxl_1024, N > 1: also time the stack gather to rank 0")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps")
    ap.add_argument("--profile-csv", default=None, help="write the per-launch table of the profiling pass here")
    args = ap.parse_args()
    rank, world, local = dist_env()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
