"""Epilogue-cost probe: same GEMM shape with increasingly heavy epilogues (CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
from probe_ops import timeit

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
print("env", {k: v for k, v in os.environ.items() if k.startswith("GDF_")})
for M, N, K in [(32768, 5120, 640), (8192, 1280, 1280), (8192, 10240, 1280), (32768, 640, 640), (8192, 8192, 8192)]:
    a, w = rb(M, K), rb(N, K)
    bias = torch.randn(N, device=dev)
    for name, geglu, res, cap, use_bias in [("plain", 0, 0, 0, 0), ("bias", 0, 0, 0, 1), ("bias+cap", 0, 0, 1, 1), ("bias+res", 0, 1, 0, 1),
                                            ("geglu", 1, 0, 0, 1), ("geglu+cap", 1, 0, 1, 1)]:
        nout = N // 2 if geglu else N
        out = torch.empty(M, nout, dtype=torch.bfloat16, device=dev)
        r = rb(M, nout) if res else None
        c = torch.empty(M, nout, dtype=torch.float16, device=dev) if cap else None
        ep = ops.make_epilogue(out=out, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, residual=r, bias=bias if use_bias else None,
                               caps=[(c, 0, nout)] if cap else ())
        ms = timeit(lambda: ops.linear(a, w, ep))
        print("M=%6d N=%6d K=%5d %-10s : %8.3f ms  %7.1f TFLOP/s" % (M, N, K, name, ms, 2.0 * M * N * K / ms / 1e9))
