"""Kernel-level timing probe at SDXL-1024 / batch-8 shapes (CUDA events, warm-up, inputs > L2 where relevant).
Run on the GPU box:  python tools/probe_ops.py > gpurun_out/probe.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
    print("device:", torch.cuda.get_device_name(0))
    # ---- linear shapes (M, N, K, geglu)
    for name, M, N, K, geglu in [("ffn1 geglu 32^2", 8192, 10240, 1280, True), ("ffn2 32^2", 8192, 1280, 5120, False),
                                 ("qkv 32^2", 8192, 3840, 1280, False), ("out-proj 32^2", 8192, 1280, 1280, False),
                                 ("ffn1 geglu 64^2", 32768, 5120, 640, True), ("ffn2 64^2", 32768, 640, 2560, False),
                                 ("qkv 64^2", 32768, 1920, 640, False), ("square 8192", 8192, 8192, 8192, False)]:
        a, w = rb(M, K), rb(N, K)
        nout = N // 2 if geglu else N
        out = torch.empty(M, nout, dtype=torch.bfloat16, device=dev)
        ep = ops.make_epilogue(out=out, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE)
        ms = timeit(lambda: ops.linear(a, w, ep))
        print("linear %-18s M=%6d N=%6d K=%5d : %8.3f ms  %7.1f TFLOP/s" % (name, M, N, K, ms,
                                                                           2.0 * M * N * K / ms / 1e9))
    # ---- conv shapes (B, H, W, Cin, Cout)
    for name, B, H, W, Cin, Cout in [("res 128^2 320", 8, 128, 128, 320, 320), ("res 64^2 640", 8, 64, 64, 640, 640),
                                     ("res 32^2 1280", 8, 32, 32, 1280, 1280), ("up 32^2 2560->1280", 8, 32, 32, 2560, 1280),
                                     ("up 128^2 960->320", 8, 128, 128, 960, 320),
                                     ("vae 1024^2 128", 2, 1024, 1024, 128, 128), ("vae 512^2 256", 8, 512, 512, 256, 256)]:
        x = rb(B, H, W, Cin)
        wp = rb(Cout, 9 * Cin)
        out = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=dev)
        ep = ops.make_epilogue(out=out)
        ms = timeit(lambda: ops.conv3x3(x, wp, ep))
        print("conv3x3 %-20s : %8.3f ms  %7.1f TFLOP/s" % (name, ms, 2.0 * B * H * W * 9 * Cin * Cout / ms / 1e9))
    # ---- attention
    for name, B, heads, N, Nk in [("self 64^2", 8, 10, 4096, 4096), ("self 32^2", 8, 20, 1024, 1024),
                                  ("cross 64^2", 8, 10, 4096, 77), ("cross 32^2", 8, 20, 1024, 77)]:
        C = heads * 64
        q, k, v = rb(B * N, C), rb(B * Nk, C), rb(B * Nk, C)
        f16 = Nk >= 128
        if f16:
            v = v.half()
        ms = timeit(lambda: ops.attention(q, k, v, B, heads, N, Nk, 0.125, v_f16=f16))
        print("attention %-12s : %8.3f ms  %7.1f TFLOP/s" % (name, ms, 4.0 * B * heads * N * Nk * 64 / ms / 1e9))
    # ---- norms
    for name, B, HW, C in [("gn 128^2 320", 8, 16384, 320), ("gn 128^2 960", 8, 16384, 960), ("gn 32^2 1280", 8, 1024, 1280),
                           ("gn vae 1024^2 128", 2, 1024 * 1024, 128)]:
        x = rb(B, HW, C)
        gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ms = timeit(lambda: ops.groupnorm(x, gm, bt, 32, 1e-5, True))
        print("groupnorm %-18s : %8.3f ms  %7.1f GB/s (3 passes alg: 2R+1W)" % (name, ms, 3.0 * x.numel() * 2 / ms / 1e6))
    x = rb(32768, 640)
    gm, bt = torch.ones(640, device=dev), torch.zeros(640, device=dev)
    ms = timeit(lambda: ops.layernorm(x, gm, bt, 1e-5))
    print("layernorm 32768x640 : %8.3f ms  %7.1f GB/s" % (ms, 2.0 * x.numel() * 2 / ms / 1e6))
    # ---- resize+concat (SDXL practical stack, B=8)
    maps = [torch.randn(8, 1024, 1280, generator=g, device=dev).half() for _ in range(2)] + \
           [torch.randn(8, 4096, 640, generator=g, device=dev).half() for _ in range(2)]
    ms = timeit(lambda: ops.resize_concat(maps, (128, 128), nhwc=True, nchw=False))
    alg = sum(m.numel() for m in maps) * 2 + 8 * 128 * 128 * 3840 * 2
    print("resize_concat nhwc B=8 3840ch : %8.3f ms  %7.1f GB/s" % (ms, alg / ms / 1e6))
    ms = timeit(lambda: ops.resize_concat(maps, (128, 128), nhwc=False, nchw=True))
    print("resize_concat nchw B=8 3840ch : %8.3f ms  %7.1f GB/s" % (ms, alg / ms / 1e6))


if __name__ == "__main__":
    main()
