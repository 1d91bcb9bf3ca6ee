"""One VAE-shaped convolution launch for an ncu capture: 1024^2, 128 -> 128 channels, 2 images, residual + bias
(the epilogue of the second conv of a VAE resnet), halo-tile mode unless GDF_CONV_HALO=0."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
x = rb(2, 1024, 1024, 128); wp = rb(128, 9 * 128)
res = rb(2 * 1024 * 1024, 128)
out = torch.empty(2 * 1024 * 1024, 128, dtype=torch.bfloat16, device=dev)
bias = torch.randn(128, device=dev)
for r in (None, res):
    for _ in range(2):
        ops.conv3x3(x, wp, ops.make_epilogue(out=out, bias=bias, residual=r))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in (None, res):
    e0.record()
    for _ in range(5):
        ops.conv3x3(x, wp, ops.make_epilogue(out=out, bias=bias, residual=r))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("conv 1024^2 128->128 x2 res=%d: %.3f ms %.0f TF/s" % (r is not None, ms, 2.0 * 2 * 1024 * 1024 * 128 * 1152 / ms / 1e9))
