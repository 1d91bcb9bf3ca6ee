"""Conv3x3 probe at VAE / UNet shapes with and without residual (env knobs: GDF_CTA_GROUP, GDF_FAST_EPI, GDF_BLOCK_N)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
from probe_ops import timeit

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
print("env", {k: v for k, v in os.environ.items() if k.startswith("GDF_")})
for name, B, H, W, Cin, Cout in [("vae 1024^2 128", 8, 1024, 1024, 128, 128), ("vae 512^2 256", 8, 512, 512, 256, 256),
                                 ("unet 128^2 320", 8, 128, 128, 320, 320), ("unet 64^2 640", 8, 64, 64, 640, 640),
                                 ("unet 32^2 1280", 8, 32, 32, 1280, 1280)]:
    x = rb(B, H, W, Cin); wp = rb(Cout, 9 * Cin)
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=dev)
    r = rb(B * H * W, Cout)
    for res in (0, 1):
        ep = ops.make_epilogue(out=out, bias=bias, residual=r if res else None)
        ms = timeit(lambda: ops.conv3x3(x, wp, ep))
        print("conv3x3 %-16s res=%d : %8.3f ms  %7.1f TFLOP/s" % (name, res, ms, 2.0 * B * H * W * 9 * Cin * Cout / ms / 1e9))
