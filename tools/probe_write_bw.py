"""Ceilings for write-dominated kernels on this GPU: fill (write only), copy (read + write), strided slab copy."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
a = torch.empty(8, 16384, 3840, dtype=torch.float16, device="cuda")
b = torch.empty_like(a)
by = a.numel() * 2
us = t(lambda: a.zero_()); print("fill 1 GB (write only): %.1f us %.0f GB/s" % (us, by / us * 1e-3))
us = t(lambda: b.copy_(a)); print("copy 1 GB (read + write): %.1f us %.0f GB/s total" % (us, 2 * by / us * 1e-3))
x = torch.randn(8, 16384, 1280, device="cuda").half()
us = t(lambda: a[:, :, :1280].copy_(x)); print("slab copy 1280 of 3840 ch: %.1f us %.0f GB/s total" % (us, 2 * x.numel() * 2 / us * 1e-3))
x4 = torch.randn(8, 1024, 3840, device="cuda").half()
us = t(lambda: torch.nn.functional.interpolate(x4.view(8, 32, 32, 3840).permute(0, 3, 1, 2), size=(128, 128), mode="bilinear"))
print("F.interpolate 32->128 x 3840 ch (NCHW view of NHWC): %.1f us" % us)
