"""Time the tcgen05 self-attention kernel at the two SDXL-1024 shapes and check it against SDPA (fp32).
Variants are selected per process: GDF_FA_V1=1 (old two-pass kernel), GDF_FA_POLY={0,4,3,2}."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

tag = "v1" if os.environ.get("GDF_FA_V1") == "1" else "v2/poly" + os.environ.get("GDF_FA_POLY", "0")
g = torch.Generator(device="cuda").manual_seed(0)
for B, heads, N in [(8, 20, 1024), (8, 10, 4096), (2, 10, 576)]:
    C = heads * 64
    qkv = torch.randn(B * N, 3 * C, generator=g, device="cuda").to(torch.bfloat16)
    q, k = qkv[:, :C], qkv[:, C:2 * C]
    v = qkv[:, 2 * C:].half().contiguous()
    for _ in range(3):
        o = ops.attention(q, k, v, B, heads, N, N, 0.125, v_f16=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 30
    e0.record()
    for _ in range(iters):
        o = ops.attention(q, k, v, B, heads, N, N, 0.125, v_f16=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 4.0 * B * heads * N * N * 64
    # accuracy on the first batch element
    qf = q[:N].float().reshape(1, N, heads, 64).transpose(1, 2)
    kf = k[:N].float().reshape(1, N, heads, 64).transpose(1, 2)
    vf = v[:N].float().reshape(1, N, heads, 64).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(N, C)
    got = o[:N].float()
    err = (got - want).abs().max().item() / want.abs().max().item()
    cos = F.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    print("%s attn B%d h%d N%d: %.1f us  %.1f TF/s  maxerr/absmax %.2e cos %.6f" % (tag, B, heads, N, us, fl / us * 1e-6, err, cos))
