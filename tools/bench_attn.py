"""Time the persistent tcgen05 attention kernel at the SDXL-1024 shapes (self: N = 1024 / 4096, text cross: Nk = 77),
the PixArt (d = 72) and Flux (d = 128) shapes, and check it against SDPA (fp32).
Variants are selected per process: GDF_ATTN_TC=0 (round-1 kernels), GDF_FA_POLY8={0,2,3,4} (share of the exponentials
evaluated by the FMA-pipe polynomial, in eighths)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

tag = "old" if os.environ.get("GDF_ATTN_TC") == "0" else "tc/split=%s/pp=%s/poly8=%s/lmode=%s" % (
    os.environ.get("GDF_FA_SPLIT", "1"), os.environ.get("GDF_FA_PP", "1"), os.environ.get("GDF_FA_POLY8", "0"),
    os.environ.get("GDF_FA_LMODE", "3"))
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(8, 20, 1024, 1024, 64, True), (8, 10, 4096, 4096, 64, True), (2, 10, 576, 576, 64, True),
          (8, 20, 1024, 77, 64, False), (8, 10, 4096, 77, 64, False)]
if os.environ.get("BENCH_ATTN_ONLY"):      # one shape (ncu captures): index into the list above
    shapes = [shapes[int(os.environ["BENCH_ATTN_ONLY"])]]
if os.environ.get("BENCH_ATTN_ALL"):
    shapes += [(8, 16, 4096, 4096, 72, False), (8, 16, 4096, 300, 72, False), (1, 24, 4608, 4608, 128, False),
               (8, 8, 4096, 4096, 40, False)]
for B, heads, N, Nk, D, f16 in shapes:
    C = heads * D
    q = torch.randn(B * N, C, generator=g, device="cuda").to(torch.bfloat16)
    kv = torch.randn(B * Nk, 2 * C, generator=g, device="cuda").to(torch.bfloat16)
    k = kv[:, :C]
    v = kv[:, C:].half().contiguous() if f16 else kv[:, C:]
    run = lambda: ops.attention(q, k, v, B, heads, N, Nk, D ** -0.5, head_dim=D, v_f16=f16)
    for _ in range(3):
        o = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    e0.record()
    for _ in range(iters):
        o = run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 4.0 * B * heads * N * Nk * D
    qf = q[:N].float().reshape(1, N, heads, D).transpose(1, 2)
    kf = k[:Nk].float().reshape(1, Nk, heads, D).transpose(1, 2)
    vf = v[:Nk].float().reshape(1, Nk, heads, D).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(N, C)
    got = o[:N].float()
    err = (got - want).abs().max().item() / want.abs().max().item()
    cos = F.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    print("%s attn B%d h%d Nq%d Nk%d d%d: %.1f us  %.1f TF/s  maxerr/absmax %.2e cos %.6f"
          % (tag, B, heads, N, Nk, D, us, fl / us * 1e-6, err, cos), flush=True)
