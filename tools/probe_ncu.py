"""A handful of representative launches of the SDXL-1024 step for one `ncu --set full` capture:
  python tools/probe_ncu.py            (run under ncu -k regex:gemm_tcgen05|attention64_tcgen05|groupnorm)
Order of launches (2 warm + 1 measured each is up to ncu's -s/-c): vae conv N=128, out-proj, FFN-out, GEGLU,
QKV, self-attention 1024, self-attention 4096, groupnorm 1024^2 x128."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)

# 1. VAE conv 1024^2 128 -> 128 (2 images)
x = rb(2, 1024, 1024, 128); wp = rb(128, 9 * 128)
out = torch.empty(2 * 1024 * 1024, 128, dtype=torch.bfloat16, device=dev)
ops.conv3x3(x, wp, ops.make_epilogue(out=out, bias=torch.randn(128, device=dev)))
# 2-5. linears
for M, N, K, geglu, res, ncap in [(8192, 1280, 1280, 0, 1, 0), (8192, 1280, 5120, 0, 1, 1), (8192, 10240, 1280, 1, 0, 1),
                                  (8192, 3840, 1280, 0, 0, 3)]:
    a, w = rb(M, K), rb(N, K)
    nout = N // 2 if geglu else N
    o = torch.empty(M, nout, dtype=torch.bfloat16, device=dev)
    r = rb(M, nout) if res else None
    caps = []
    if ncap == 1:
        caps = [(torch.empty(M, nout, dtype=torch.float16, device=dev), 0, nout)]
    elif ncap == 3:
        caps = [(torch.empty(M, N // 3, dtype=torch.float16, device=dev), i * (N // 3), (i + 1) * (N // 3)) for i in range(3)]
    ep = ops.make_epilogue(out=o, bias=torch.randn(N, device=dev), act=ops.ACT_GEGLU if geglu else ops.ACT_NONE,
                           residual=r, caps=caps)
    ops.linear(a, w, ep, block_n=256 if geglu else 0)
# 6-7. self attention
for B, heads, N in [(8, 20, 1024), (8, 10, 4096)]:
    C = heads * 64
    q, k = rb(B * N, C), rb(B * N, C)
    v = rb(B * N, C).half()
    ops.attention(q, k, v, B, heads, N, N, 0.125, v_f16=True)
# 8. groupnorm VAE level 0
xg = rb(2, 1024 * 1024, 128)
ops.groupnorm(xg, torch.ones(128, device=dev), torch.zeros(128, device=dev), 32, 1e-6, True)
torch.cuda.synchronize()
print("done")
