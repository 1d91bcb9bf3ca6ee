"""One launch each of the K = 1280 projection shapes for an `ncu --set full` capture: is the N = K = 1280 GEMM bound by
operand delivery from L2 (lts throughput) rather than by the tensor pipe?  ("What is next" 2)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import ops

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rb = lambda *s: torch.randn(*s, generator=g, device=dev).to(torch.bfloat16)
cases = [("out-proj bn=160", 8192, 1280, 1280, 0), ("out-proj bn=256", 8192, 1280, 1280, 256), ("ffn2 bn=160", 8192, 1280, 5120, 0)]
for name, M, N, K, bn in cases:
    a, w, r = rb(M, K), rb(N, K), rb(M, N)
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    ep = ops.make_epilogue(out=out, residual=r)
    for _ in range(3):
        ops.linear(a, w, ep, block_n=bn)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    ops.linear(a, w, ep, block_n=bn)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(name, "done")
