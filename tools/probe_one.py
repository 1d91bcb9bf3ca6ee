"""Single-shape GEMM launch for ncu: python probe_one.py M N K geglu res cap iters"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
M, N, K, geglu, res, cap, iters = [int(x) for x in sys.argv[1:8]]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
a, w = rb(M, K), rb(N, K)
nout = N // 2 if geglu else N
out = torch.empty(M, nout, dtype=torch.bfloat16, device=dev)
r = rb(M, nout) if res else None
c = torch.empty(M, nout, dtype=torch.float16, device=dev) if cap else None
ep = ops.make_epilogue(out=out, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, residual=r, caps=[(c, 0, nout)] if cap else ())
for _ in range(iters):
    ops.linear(a, w, ep)
torch.cuda.synchronize()
print("done")
