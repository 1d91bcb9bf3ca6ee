"""Timeline of CTA 0 of the persistent attention kernel (gdf_debug_attention_trace): role events stamped with the SM
clock, printed per role in time order (clock deltas in cycles).   python tools/attn_trace.py B heads Nq Nk [D]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from generic_diffusion_feature_b200 import _lib, ops

NAMES = {1: "Q0 load", 2: "Q1 load", 3: "K load", 4: "V load", 10: "S0 issue", 11: "S1 issue", 12: "PV0 issue",
         13: "PV1 issue", 20: "sm0 S seen", 21: "sm1 S seen", 22: "sm0 max done", 23: "sm1 max done", 24: "sm0 P free",
         25: "sm1 P free", 26: "sm0 P published", 27: "sm1 P published", 28: "sm0 out start", 29: "sm1 out start",
         30: "sm0 out done", 31: "sm1 out done"}
B, heads, Nq, Nk = [int(x) for x in sys.argv[1:5]]
D = int(sys.argv[5]) if len(sys.argv) > 5 else 64
f16 = Nk >= 128
g = torch.Generator(device="cuda").manual_seed(0)
C = heads * D
q = torch.randn(B * Nq, C, generator=g, device="cuda").to(torch.bfloat16)
kv = torch.randn(B * Nk, 2 * C, generator=g, device="cuda").to(torch.bfloat16)
k = kv[:, :C]
v = kv[:, C:].half().contiguous() if f16 else kv[:, C:]
lib = _lib.load()
for _ in range(3):
    ops.attention(q, k, v, B, heads, Nq, Nk, D ** -0.5, head_dim=D, v_f16=f16)
cap = 1 << 15
buf = torch.zeros(cap, dtype=torch.int64, device="cuda")
_lib.check(lib.gdf_debug_attention_trace(_lib.ptr(buf), cap))
ops.attention(q, k, v, B, heads, Nq, Nk, D ** -0.5, head_dim=D, v_f16=f16)
torch.cuda.synchronize()
_lib.check(lib.gdf_debug_attention_trace(None, 0))
h = buf.cpu().numpy().astype("uint64")
n = int(h[0])
ev = sorted(((int(x) & 0xFFFFFFFFFF, int(x) >> 56, (int(x) >> 40) & 0xFFFF) for x in h[1:1 + min(n, cap - 1)]))
t0 = ev[0][0]
print("# %d events; columns: clock-from-start, delta, event, tile/item index" % n)
prev = t0
limit = int(os.environ.get("TRACE_LIMIT", "400"))
for t, kind, gidx in ev[:limit]:
    print("%9d %6d  %-16s %d" % (t - t0, t - prev, NAMES.get(kind, str(kind)), gidx))
    prev = t
