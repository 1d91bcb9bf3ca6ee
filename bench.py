#!/usr/bin/env python
"""Benchmark of the extraction hot path: SDXL 1024x1024 full-activation feature extraction, images/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one batch of synthetic images: VAE encode + posterior sample +
q_sample (t=50) + one SDXL UNet forward with all 472 non-`map` activations of config_xl_full captured at native
resolution into the fp16 arena (BASELINE.json configs[1]). Weights are random-init by parameter name, images /
conditioning / noise are synthetic (no network for checkpoints or datasets).

Rank 0 prints ONE JSON line (contract in the task statement): `value` = whole-job images/s with inputs resident
in HBM, `e2e` = the same through FeatureExtractor.extract with pinned-host images (H2D inside the timed region)
and a D2H read of the `unet-out` map, `roofline` for the dominant kernel (the tcgen05 GEMM / implicit-GEMM conv)
from a CUDA-event profiling pass, `cpu_baseline` = the CPU oracle port timed on the host cores (N=1 only).
`--impl reference` times the CPU oracle port (the reference's diffusers path cannot be imported: diffusers is not
installed and cannot be) on the same config.
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

METRIC = "sdxl_1024_full_activation_extraction_images_per_s"
UNIT = "images/s"
FLOP_PER_IMAGE = 11.640e12      # UNet 6.761 + VAE encoder 4.879 TFLOP / image
IMG = 1024


def read_gemm_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu launch list of this same command
    (profiles/r01_gemm_traffic.json, written by tools/ncu_launch_summary.py --traffic-json); None if absent."""
    p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:  # noqa: BLE001
            return None
    return None


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1371.9), d.get("bf16_tflops", 1668.4), d.get("hbm_gbs", 6540.2), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


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


def full_xl_layer():
    from generic_diffusion_feature_b200.components.feature_extractor import _unet_feature_ids
    from generic_diffusion_feature_b200.components.models import UNET_CONFIGS
    return {i: True for i in _unet_feature_ids(UNET_CONFIGS["xl"])}   # == non-map ids of config_xl_full.json


def cpu_oracle_images_per_s(sd, steps, warmup, batch=1):
    """CPU port of the reference path (oracle) on the host cores: `batch` images per step at 1024^2, full set."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import O, build_oracle, make_inputs
    from generic_diffusion_feature_b200.components.models import UNET_CONFIGS, VAE_CONFIGS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae = build_oracle(UNET_CONFIGS["xl"], VAE_CONFIGS["xl"], sd)
    layer = full_xl_layer()
    store = O.FeatureStore(layer)
    O.attach_gatherers(unet, store)
    image, ctx, pooled, ev, eq = make_inputs(batch, IMG, 2048, 1280)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.extract("xl", unet, vae, store, image, ctx, pooled, ev, eq, t=50, img_size=IMG)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, cores, total / len(times)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from generic_diffusion_feature_b200.components.models import synthetic_state_dict
    sd = synthetic_state_dict("xl", "cpu")
    ips, cores, sec = cpu_oracle_images_per_s(sd, args.steps, args.warmup, batch=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SDXL UNet 1024x1024 full activation set (472 maps), VAE encode + q_sample t=50",
                   "batch_per_step": 1, "parallelism": "host threads"},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 image 1024x1024 per step through the CPU oracle (torch fp32 restatement of "
                                   "the reference path; diffusers itself is not installable offline)"},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, rank, world, local):
    import torch.distributed as dist
    from generic_diffusion_feature_b200 import _lib
    from generic_diffusion_feature_b200.components.models import get_diffusion_model
    from generic_diffusion_feature_b200.diffusion_feature import FeatureExtractor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the extraction path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line of the contract
        dist.init_process_group("nccl", device_id=torch.device(dev))
    B = args.batch
    pipe = get_diffusion_model("xl", "float16", device=dev, weight_device=dev, synthetic=True)
    fe = FeatureExtractor(full_xl_layer(), "xl", dev, img_size=IMG, external_model=pipe)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    images = torch.rand(B, 3, IMG, IMG, generator=g, device=dev) * 2 - 1
    L = IMG // 8
    noise = (torch.randn(B, 4, L, L, generator=g, device=dev), torch.randn(B, 4, L, L, generator=g, device=dev))
    prompts = tuple(None if p is None else p.to(dev) for p in fe.encode_prompt(""))
    images_host = images.cpu().pin_memory()

    def step_device():
        return fe.extract(prompts, B, images, image_type="tensors", t=50, noise=noise)

    def step_e2e():
        feats = fe.extract(prompts, B, images_host, image_type="tensors", t=50, noise=noise)
        return feats["unet-out"].cpu()       # D2H read of the step's result (also syncs the step)

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

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.cuda_profiler:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: capture the timed steps only
    ms = timed(step_device, args.steps)
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms / 1e3)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel: CUDA-event profiling pass on the launching stream
    lib = pipe.lib
    roof = None
    kinds = ["tcgen05_gemm_conv", "attention", "groupnorm", "layernorm", "other"]
    if rank == 0:
        _lib.check(lib.gdf_profile(pipe.handle, 1))
        for _ in range(2):
            step_device()
        torch.cuda.synchronize()
        msv = (ctypes.c_float * 5)()
        flv = (ctypes.c_double * 5)()
        lnv = (ctypes.c_int * 5)()
        _lib.check(lib.gdf_profile_read(pipe.handle, msv, flv, lnv))
        if args.profile_csv:
            _lib.check(lib.gdf_profile_dump(pipe.handle, args.profile_csv.encode()))
        _lib.check(lib.gdf_profile(pipe.handle, 0))
        sustained, burst, hbm, how = read_peaks()
        gemm_tflops = flv[0] / (msv[0] * 1e-3) / 1e12 if msv[0] > 0 else 0.0
        tot_ms = sum(msv)
        traffic = read_gemm_traffic()
        roof = {"bound": "tensor", "achieved": gemm_tflops, "peak": sustained, "unit": "TFLOP/s",
                "frac": gemm_tflops / sustained,
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "traffic_source": traffic["source"] if traffic else None,
                "algorithmic_flop_per_launch": flv[0] / max(lnv[0], 1),
                "peak_source": how + " (sustained bf16)",
                "kernel": "gemm_tcgen05_kernel",
                "avg_launch_us": 1e3 * msv[0] / max(lnv[0], 1),
                "share_of_step": msv[0] / tot_ms if tot_ms > 0 else None,
                "per_kind_ms_per_step": {k: msv[i] / 2 for i, k in enumerate(kinds)},
                "per_kind_launches_per_step": {k: lnv[i] // 2 for i, k in enumerate(kinds)},
                "whole_path_tflops": FLOP_PER_IMAGE * value / world / 1e12,
                "whole_path_frac": FLOP_PER_IMAGE * value / world / 1e12 / sustained}

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sd = {k: v for k, v in _cpu_state_dict(dev).items()}
            ips, cores, sec = cpu_oracle_images_per_s(sd, 1, 0, batch=1)
            cpu_base = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "1 image 1024x1024, full activation set, CPU oracle (torch fp32), %.1f s" % sec}
        except Exception as ex:  # noqa: BLE001
            cpu_base = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                        "sample": "failed: %r" % (ex,)}

    if rank == 0:
        launches_per_step = fe._plan.launches + 3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SDXL UNet 1024x1024 full activation set (472 maps, fp16 arena), VAE encode + "
                                   "q_sample t=50", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": "dp%d (replicated weights, images sharded, no data-path collective)" % world,
                       "l2": "working set per step (5 GB weights + 17 GB arena) >> 126 MB L2, no flush needed",
                       "arena_gb": fe._plan.arena_bytes / 1e9, "workspace_gb": fe._plan.workspace_bytes / 1e9},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": images_host.numel() * 4,
                    "d2h_bytes_per_step": B * 4 * L * L * 2, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roof, "cpu_baseline": cpu_base,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def _cpu_state_dict(dev):
    from generic_diffusion_feature_b200.components.models import synthetic_state_dict
    sd = synthetic_state_dict("xl", dev)
    return {k: v.cpu() for k, v in sd.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
