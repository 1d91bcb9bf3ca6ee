"""Two self-attention launches (SDXL-1024 shapes) for one `ncu --set full --import-source on` capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
for B, heads, N in [(8, 20, 1024), (8, 10, 4096)]:
    C = heads * 64
    qkv = torch.randn(B * N, 3 * C, generator=g, device="cuda").to(torch.bfloat16)
    v = qkv[:, 2 * C:].half().contiguous()
    o = ops.attention(qkv[:, :C], qkv[:, C:2 * C], v, B, heads, N, N, 0.125, v_f16=True)
torch.cuda.synchronize()
print("done")
