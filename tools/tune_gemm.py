"""Sweep block_n x cta_group on the GEMM / conv shapes of the SDXL-1024 B=8 step (env knobs read per launch build)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
from probe_ops import timeit

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
BNS = [96, 112, 128, 144, 160, 176, 192, 208, 224, 240, 256]


def sweep(tag, fn, flops, bns=BNS):
    res = []
    for cg in (1, 2):
        for bn in bns:
            os.environ["GDF_CTA_GROUP"] = str(cg)
            os.environ["GDF_BLOCK_N"] = str(bn)
            try:
                ms = timeit(fn, iters=8, warm=2)
            except Exception as ex:
                print("  fail", cg, bn, ex)
                continue
            res.append((ms, cg, bn))
    res.sort()
    best = res[0]
    line = " ".join("cg%d/bn%d=%.1fus" % (c, b, m * 1e3) for m, c, b in res[:6])
    worst = res[-1]
    print("%-34s best cg%d bn%3d %7.1f us %7.1f TF | %s | worst cg%d/bn%d=%.1fus" % (tag, best[1], best[2], best[0] * 1e3, flops / best[0] / 1e9, line, worst[1], worst[2], worst[0] * 1e3), flush=True)


lin = [(8192, 1280, 1280, 1, 0), (8192, 1280, 1280, 0, 1), (8192, 1280, 5120, 1, 1), (8192, 3840, 1280, 0, 3), (8192, 1280, 2560, 0, 0),
       (32768, 640, 640, 1, 0), (32768, 640, 640, 0, 1), (32768, 640, 2560, 1, 1), (32768, 1920, 640, 0, 3), (616, 2560, 2048, 0, 0), (616, 1280, 2048, 0, 0),
       (131072, 320, 320, 0, 0)]
for M, N, K, res, ncap in lin:
    a, w = rb(M, K), rb(N, K)
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    r = rb(M, N) if res else None
    caps = []
    if ncap == 1:
        caps = [(torch.empty(M, N, dtype=torch.float16, device=dev), 0, N)]
    elif ncap == 3:
        caps = [(torch.empty(M, N // 3, dtype=torch.float16, device=dev), i * (N // 3), (i + 1) * (N // 3)) for i in range(3)]

    def fn():
        ep = ops.make_epilogue(out=out, bias=bias, residual=r, caps=caps)
        ops.linear(a, w, ep)
    sweep("lin M=%d N=%d K=%d res=%d cap=%d" % (M, N, K, res, ncap), fn, 2.0 * M * N * K)

if os.environ.get('TUNE_LIN_ONLY'):
    sys.exit(0)
for B, H, W, Cin, Cout, res in [(8, 32, 32, 1280, 1280, 1), (8, 32, 32, 2560, 1280, 0), (8, 64, 64, 640, 640, 1), (8, 64, 64, 1280, 640, 0), (8, 64, 64, 1920, 640, 0),
                                (8, 128, 128, 320, 320, 1), (8, 128, 128, 640, 320, 0), (8, 128, 128, 960, 320, 0), (8, 1024, 1024, 128, 128, 0), (8, 512, 512, 128, 256, 0), (8, 512, 512, 256, 256, 0)]:
    x = rb(B, H, W, Cin); wp = rb(Cout, 9 * Cin)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=dev)
    r = rb(B * H * W, Cout) if res else None

    def fn():
        ep = ops.make_epilogue(out=out, bias=bias, residual=r)
        ops.conv3x3(x, wp, ep)
    bns = [b for b in BNS if b <= max(Cout, 96)]
    sweep("conv %dx%d Cin=%d Cout=%d res=%d" % (H, W, Cin, Cout, res), fn, 2.0 * B * H * W * 9 * Cin * Cout, bns)
