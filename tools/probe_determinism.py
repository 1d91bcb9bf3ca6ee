"""Which ops are bit-reproducible from launch to launch? (GDF_DETERMINISTIC=1 expected: all of them)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

g = torch.Generator(device="cuda").manual_seed(3)
rb = lambda *s: torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)

def rep(name, fn, n=6):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    bad = sum(1 for o in outs[1:] if not all(torch.equal(a, b) for a, b in zip(o, outs[0])))
    print("%-60s %s" % (name, "bit-identical" if bad == 0 else "DIFFERS in %d of %d repeats" % (bad, n - 1)), flush=True)

img = torch.randn(2, 3, 256, 256, generator=g, device="cuda")
w_in = torch.randn(128, 3, 3, 3, generator=g, device="cuda") * 0.2
b_in = torch.randn(128, generator=g, device="cuda")
rep("conv_in fused (no stats)", lambda: (ops.conv_in_fused(img, w_in, b_in)[0],))
rep("conv_in fused (+ GroupNorm sums, atomics)", lambda: ops.conv_in_fused(img, w_in, b_in, gn_groups=32))
for (B, H, Cin, Cout, res) in [(2, 128, 128, 128, False), (2, 128, 128, 128, True), (2, 64, 256, 256, True), (1, 64, 320, 320, True)]:
    x = rb(B, H, H, Cin); wp = rb(Cout, 9 * Cin) * 0.05
    r = rb(B * H * H, Cout) if res else None
    bias = torch.randn(Cout, generator=g, device="cuda")
    def f():
        out = torch.zeros(B * H * H, Cout, dtype=torch.bfloat16, device="cuda")
        ops.conv3x3(x, wp, ops.make_epilogue(out=out, bias=bias, residual=r))
        return (out,)
    rep("conv3x3 %dx%d %d->%d res=%d" % (H, H, Cin, Cout, res), f)
x = rb(2, 128 * 128, 128); gm = torch.randn(128, device="cuda"); bt = torch.randn(128, device="cuda")
rep("groupnorm 128x128x128", lambda: (ops.groupnorm(x, gm, bt, 32, 1e-6, True),))
x2 = rb(2, 32 * 32, 320); gm2 = torch.randn(320, device="cuda"); bt2 = torch.randn(320, device="cuda")
rep("groupnorm 32x32x320", lambda: (ops.groupnorm(x2, gm2, bt2, 32, 1e-5, True),))
a, w = rb(4096, 640), rb(640, 640) * 0.05
r = rb(4096, 640)
def f2():
    out = torch.zeros(4096, 640, dtype=torch.bfloat16, device="cuda")
    ops.linear(a, w, ops.make_epilogue(out=out, residual=r))
    return (out,)
rep("linear 4096x640x640 res", f2)
q, k, v = rb(2 * 1024, 640), rb(2 * 1024, 640), rb(2 * 1024, 640).half()
rep("attention d64 N=1024", lambda: (ops.attention(q, k, v, 2, 10, 1024, 1024, 0.125, v_f16=True),))
xl = rb(4096, 640); gl = torch.randn(640, device="cuda"); bl = torch.randn(640, device="cuda")
rep("layernorm", lambda: (ops.layernorm(xl, gl, bl, 1e-5),))
