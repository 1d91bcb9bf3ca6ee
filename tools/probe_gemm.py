"""GEMM-only timing probe (tuning knobs via env: GDF_BLOCK_N, GDF_TMA_STORE)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
from probe_ops import timeit

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
print("env", {k: v for k, v in os.environ.items() if k.startswith("GDF_")})
for name, M, N, K, geglu, res, cap in [("ffn1 geglu", 8192, 10240, 1280, True, False, True), ("ffn2", 8192, 1280, 5120, False, True, True),
                                       ("qkv", 8192, 3840, 1280, False, False, True), ("out-proj", 8192, 1280, 1280, False, True, False),
                                       ("ffn1 64^2", 32768, 5120, 640, True, False, True), ("qkv 64^2", 32768, 1920, 640, False, False, True),
                                       ("outproj 64^2", 32768, 640, 640, False, True, False)]:
    a, w = rb(M, K), rb(N, K)
    nout = N // 2 if geglu else N
    out = torch.empty(M, nout, dtype=torch.bfloat16, device=dev)
    r = rb(M, nout) if res else None
    c = torch.empty(M, nout, dtype=torch.float16, device=dev) if cap else None
    ep = ops.make_epilogue(out=out, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, residual=r,
                           caps=[(c, 0, nout)] if cap else ())
    ms = timeit(lambda: ops.linear(a, w, ep))
    print("%-14s M=%6d N=%6d K=%5d res=%d cap=%d : %8.3f ms  %7.1f TFLOP/s" % (name, M, N, K, res, cap, ms, 2.0 * M * N * K / ms / 1e9))
for name, B, H, W, Cin, Cout in [("res 32^2 1280", 8, 32, 32, 1280, 1280), ("vae 1024^2 128", 2, 1024, 1024, 128, 128)]:
    x = rb(B, H, W, Cin); wp = rb(Cout, 9 * Cin)
    out = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=dev)
    ep = ops.make_epilogue(out=out)
    ms = timeit(lambda: ops.conv3x3(x, wp, ep))
    print("conv3x3 %-20s : %8.3f ms  %7.1f TFLOP/s" % (name, ms, 2.0 * B * H * W * 9 * Cin * Cout / ms / 1e9))
