"""Per-kernel parity tests, called through the C ABI (ctypes -> libgdf_b200.so) on a B200.

Each CUDA kernel is compared with a plain PyTorch fp32 evaluation of the same op on the same seeded inputs.
Tolerances are for bf16 inputs/outputs with fp32 accumulation: relative L2 error <= 1e-2 (bf16 has 8 bits of
mantissa, rounding of the output alone is 2^-9 relative) and cosine similarity >= 0.9999.
"""
import math
import os

import numpy as np

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ops():
    from generic_diffusion_feature_b200 import ops
    return ops


def rel_err(a, b):
    a = a.float().flatten()
    b = b.float().flatten()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cos_sim(a, b):
    return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def check_close(got, want, tol=1e-2, what=""):
    assert torch.isfinite(got.float()).all(), what + ": non-finite output"
    r, c = rel_err(got, want), cos_sim(got, want)
    assert r <= tol and c >= 0.9999, "%s: rel_err %.3e cos %.6f" % (what, r, c)


def _rand_bf16(gen, *shape, scale=1.0, dev="cuda"):
    return (torch.randn(*shape, generator=gen, device=dev) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(300, 320, 640), (128, 16, 64), (1024, 1920, 640), (4096, 1280, 5120),
                                    (77 * 2, 2560, 2048), (130, 640, 2816)])
def test_linear_bias(cuda_dev, M, N, K):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    a = _rand_bf16(g, M, K)
    w = _rand_bf16(g, N, K, scale=K ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    ops.linear(a, w, ops.make_epilogue(out=out, bias=bias))
    torch.cuda.synchronize()
    check_close(out, a.float() @ w.float().T + bias, what="linear %dx%dx%d" % (M, N, K))


@pytest.mark.parametrize("M,C,N,geglu", [(1000, 640, 1920, False), (8192, 1280, 1280, False), (700, 320, 2560, True),
                                         (300, 1280, 10240, True)])
def test_linear_layernorm_fold(cuda_dev, M, C, N, geglu):
    """LayerNorm folded into the consuming projection (attention.py:497,525,566): the producer GEMM adds per-row
    (sum, sum sq) of its output (row_sums), the consumer runs on the un-normalised rows with gamma folded into the
    weight and applies rstd * (acc - mean * u) + (bias + beta.W^T) in its epilogue. Reference: F.layer_norm + matmul."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(40)
    # producer: x = a @ w0^T + b0 + res  (bf16), statistics of x into sums
    a = _rand_bf16(g, M, 320)
    w0 = _rand_bf16(g, C, 320, scale=320 ** -0.5)
    b0 = torch.randn(C, generator=g, device="cuda")
    res = _rand_bf16(g, M, C) + 0.5   # non-zero row mean
    x = torch.zeros(M, C, dtype=torch.bfloat16, device="cuda")
    sums = torch.zeros(M, 2, device="cuda")
    ops.linear(a, w0, ops.make_epilogue(out=x, bias=b0, residual=res, row_sums=sums))
    torch.cuda.synchronize()
    xf = a.float() @ w0.float().T + b0 + res.float()
    assert torch.allclose(sums[:, 0], xf.sum(1), rtol=1e-3, atol=1e-2), "row sums"
    assert torch.allclose(sums[:, 1], (xf * xf).sum(1), rtol=1e-3, atol=1e-2), "row sums of squares"
    # consumer
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device="cuda")
    beta = 0.1 * torch.randn(C, generator=g, device="cuda")
    W = torch.randn(N, C, generator=g, device="cuda") * C ** -0.5
    bias = torch.randn(N, generator=g, device="cuda")
    Wg = (W * gamma).to(torch.bfloat16)
    u = Wg.float().sum(1).contiguous()
    c = (bias + W @ beta).contiguous()
    n_out = N // 2 if geglu else N
    if geglu:   # value / gate rows interleaved per 256-column tile (128 + 128), as the executor packs them
        idx = []
        for t in range(n_out // 128):
            idx += list(range(t * 128, t * 128 + 128)) + list(range(n_out + t * 128, n_out + t * 128 + 128))
        idx = torch.tensor(idx, device="cuda")
        Wg_k, u_k, c_k = Wg[idx].contiguous(), u[idx].contiguous(), c[idx].contiguous()
    else:
        Wg_k, u_k, c_k = Wg, u, c
    out = torch.zeros(M, n_out, dtype=torch.bfloat16, device="cuda")
    ep = ops.make_epilogue(out=out, bias=c_k, act=ops.ACT_GEGLU if geglu else ops.ACT_NONE, ln_sums=sums, ln_u=u_k,
                           ln_eps=1e-5)
    ops.linear(x, Wg_k, ep, block_n=256 if geglu else 0)
    torch.cuda.synchronize()
    y = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5) @ W.T + bias
    want = y[:, :n_out] * F.gelu(y[:, n_out:]) if geglu else y
    check_close(out, want, what="LN-folded linear M%d C%d N%d geglu=%d" % (M, C, N, geglu))


def test_linear_full_epilogue(cuda_dev):
    """bias + per-sample row bias + residual + out_scale + pre/post captures + second destination + fp32 out."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    B, rows, N, K = 3, 200, 640, 320
    M = B * rows
    a = _rand_bf16(g, M, K)
    w = _rand_bf16(g, N, K, scale=K ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    rbb = torch.randn(B, N, generator=g, device="cuda")
    res = _rand_bf16(g, M, N)
    out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    big = torch.zeros(M, N + 320, dtype=torch.bfloat16, device="cuda")  # concat buffer: write into [:, 320:]
    out2 = big[:, 320:]
    of32 = torch.zeros(M, N, device="cuda")
    cap_pre = torch.zeros(M, N, dtype=torch.float16, device="cuda")
    cap_a = torch.zeros(M, 320, dtype=torch.float16, device="cuda")
    cap_b = torch.zeros(M, 320, dtype=torch.float16, device="cuda")
    ep = ops.make_epilogue(out=out, bias=bias, row_batch_bias=rbb, rows_per_batch=rows, residual=res, out_scale=0.5,
                           out2=out2, out_f32=of32, cap_pre=cap_pre, caps=[(cap_a, 0, 320), (cap_b, 320, 640)])
    ops.linear(a, w, ep)
    torch.cuda.synchronize()
    pre = a.float() @ w.float().T + bias + rbb.repeat_interleave(rows, 0)
    want = (pre + res.float()) * 0.5
    check_close(cap_pre, pre, what="cap_pre")
    check_close(out, want, what="out")
    check_close(out2, want, what="out2 (strided)")
    check_close(of32, want, tol=2e-3, what="out_f32")
    check_close(cap_a, want[:, :320], what="cap seg 0")
    check_close(cap_b, want[:, 320:], what="cap seg 1")
    assert big[:, :320].abs().max().item() == 0, "strided destination wrote outside its slice"


def test_linear_geglu(cuda_dev):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    M, C = 520, 640
    inner = 4 * C
    a = _rand_bf16(g, M, C)
    w = _rand_bf16(g, 2 * inner, C, scale=C ** -0.5)
    bias = torch.randn(2 * inner, generator=g, device="cuda") * 0.1
    bn = 256
    half = bn // 2
    # interleave value / gate rows per 256-column tile (what the weight packer does for GEGLU)
    idx = []
    for t in range(inner // half):
        idx += list(range(t * half, (t + 1) * half)) + list(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx, device="cuda")
    out = torch.zeros(M, inner, dtype=torch.bfloat16, device="cuda")
    cap = torch.zeros(M, inner, dtype=torch.float16, device="cuda")
    ep = ops.make_epilogue(out=out, bias=bias[idx].contiguous(), act=ops.ACT_GEGLU, caps=[(cap, 0, inner)])
    ops.linear(a, w[idx].contiguous(), ep, block_n=bn)
    torch.cuda.synchronize()
    proj = a.float() @ w.float().T + bias
    want = proj[:, :inner] * F.gelu(proj[:, inner:])
    check_close(out, want, what="geglu out")
    check_close(cap, want, what="geglu capture")


@pytest.mark.parametrize("M,N,K,res,bn", [(8192, 1280, 1280, True, 0), (8192, 1280, 5120, True, 0), (8192, 3840, 1280, False, 0),
                                          (8192, 1280, 1280, False, 256), (9000, 1184, 1288, True, 0)])
def test_linear_k_split_tail(cuda_dev, M, N, K, res, bn):
    """K-split of the last partial wave (GemmParams::sk_*, ops_gemm.cu plan_k_split): the tiles that do not fill a
    whole wave of the 74 resident CTA pairs are split along K into pieces that run side by side; partial fp32
    accumulators meet in the workspace and the last piece's warps sum them in index order. Same product as the unsplit
    launch up to fp32 summation order (bf16 outputs: equal except where a rounding boundary is crossed), bit-identical
    run to run, counters left at zero."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(77)
    a = _rand_bf16(g, M, K)
    w = _rand_bf16(g, N, K, scale=K ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    r = _rand_bf16(g, M, N) if res else None
    ws = ops.k_split_workspace()
    outs = []
    for ks in (None, ws, ws):
        out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
        cap = torch.zeros(M, N, dtype=torch.float16, device="cuda")
        ops.linear(a, w, ops.make_epilogue(out=out, bias=bias, residual=r, caps=[(cap, 0, N)], k_split=ks), block_n=bn)
        torch.cuda.synchronize()
        outs.append((out, cap))
    want = a.float() @ w.float().T + bias + (r.float() if res else 0)
    for out, cap in outs:
        check_close(out, want, what="k-split linear %dx%dx%d" % (M, N, K))
        check_close(cap, want, what="k-split capture")
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[1][1], outs[2][1])      # deterministic
    assert int(ws[1].abs().sum().item()) == 0                                                # counters self-clean
    diff = (outs[0][0].float() - outs[1][0].float()).abs().max().item()
    assert diff <= 2 ** -6 * want.abs().max().item()          # one bf16 ulp at the top of the range


def test_linear_k_split_geglu(cuda_dev):
    """K-split with the GEGLU epilogue: value and gate halves of a tile are summed by the warp that consumes them."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(78)
    M, C = 8192, 1280
    inner = 1280 + 256        # 12 n-tiles of 256 accumulator columns x 32 row units = 384 tiles: 5 waves + 14
    a = _rand_bf16(g, M, C)
    w = _rand_bf16(g, 2 * inner, C, scale=C ** -0.5)
    bias = torch.randn(2 * inner, generator=g, device="cuda") * 0.1
    half = 128
    idx = []
    for t in range(inner // half):
        idx += list(range(t * half, (t + 1) * half)) + list(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx, device="cuda")
    ws = ops.k_split_workspace()
    res = []
    for ks in (None, ws):
        out = torch.zeros(M, inner, dtype=torch.bfloat16, device="cuda")
        ep = ops.make_epilogue(out=out, bias=bias[idx].contiguous(), act=ops.ACT_GEGLU, k_split=ks)
        ops.linear(a, w[idx].contiguous(), ep, block_n=256)
        torch.cuda.synchronize()
        res.append(out)
    proj = a.float() @ w.float().T + bias
    want = proj[:, :inner] * F.gelu(proj[:, inner:])
    check_close(res[0], want, what="geglu unsplit")
    check_close(res[1], want, what="geglu k-split")
    assert int(ws[1].abs().sum().item()) == 0


def test_linear_batched_and_rowbias(cuda_dev):
    """QK^T-style batched product with alpha, and a transposed product with a row (M) bias."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(4)
    Bt, N, D = 3, 384, 512
    q = _rand_bf16(g, Bt, N, D)
    k = _rand_bf16(g, Bt, N, D)
    s = torch.zeros(Bt, N, N, dtype=torch.bfloat16, device="cuda")
    ep = ops.make_epilogue(out=s, alpha=D ** -0.5, out_batch_stride=N * N)
    ops.linear(q, k, ep, batch=Bt, a_batch_stride=N * D, w_batch_stride=N * D)
    torch.cuda.synchronize()
    check_close(s, torch.einsum("bnd,bmd->bnm", q.float(), k.float()) * D ** -0.5, what="batched QK^T")
    # V^T = Wv X^T + b[:, None]
    x = _rand_bf16(g, N, D)
    wv = _rand_bf16(g, D, D, scale=D ** -0.5)
    bv = torch.randn(D, generator=g, device="cuda")
    vt = torch.zeros(D, N, dtype=torch.bfloat16, device="cuda")
    ops.linear(wv, x, ops.make_epilogue(out=vt, bias_m=bv))
    torch.cuda.synchronize()
    check_close(vt, wv.float() @ x.float().T + bv[:, None], what="transposed product with row bias")


@pytest.mark.parametrize("B,H,W,Cin,Cout,stride,pad_lo", [
    (2, 32, 32, 64, 128, 1, 1), (1, 128, 128, 128, 64, 1, 1), (3, 16, 16, 320, 320, 1, 1),
    (2, 8, 8, 128, 256, 1, 1), (2, 64, 64, 64, 4, 1, 1), (2, 64, 64, 128, 128, 2, 1),
    (2, 64, 64, 128, 128, 2, 0), (1, 256, 256, 64, 64, 2, 0), (5, 4, 4, 64, 64, 1, 1)])
def test_conv3x3(cuda_dev, B, H, W, Cin, Cout, stride, pad_lo):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = _rand_bf16(g, B, Cin, H, W)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, device="cuda") * (9 * Cin) ** -0.5
    bias = torch.randn(Cout, generator=g, device="cuda")
    wq = w.to(torch.bfloat16).float()
    if stride == 1:
        want = F.conv2d(x.float(), wq, bias, padding=1)
    elif pad_lo == 1:
        want = F.conv2d(x.float(), wq, bias, stride=2, padding=1)
    else:
        want = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), wq, bias, stride=2)
    Ho, Wo = want.shape[-2:]
    wp = ops.pack_conv_weight(w)
    npad = wp.shape[0]
    bias_p = torch.zeros(npad, device="cuda")
    bias_p[:Cout] = bias
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = torch.zeros(B * Ho * Wo, Cout, dtype=torch.bfloat16, device="cuda")
    ops.conv3x3(x_nhwc, wp, ops.make_epilogue(out=out, bias=bias_p, n_out=Cout), stride=stride, pad_lo=pad_lo)
    torch.cuda.synchronize()
    got = out.view(B, Ho, Wo, Cout).permute(0, 3, 1, 2)
    check_close(got, want, what="conv3x3 B%d %dx%d %d->%d s%d p%d" % (B, H, W, Cin, Cout, stride, pad_lo))


@pytest.mark.parametrize("B,H,Cin,Cout", [(1, 128, 320, 320), (2, 64, 640, 640), (1, 128, 960, 320)])
def test_conv3x3_resnet_epilogue_long_k(cuda_dev, B, H, Cin, Cout):
    """The second conv of a UNet resnet at its real shapes (K = 9 Cin >= 45 k-blocks: the deep-ring / single staging
    round configuration): per-sample row bias (time embedding), residual, `increment` capture before the residual,
    `out` capture after it, second bf16 destination (skip-concat slice) - every destination against torch."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    x = _rand_bf16(g, B, Cin, H, H)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, device="cuda") * (9 * Cin) ** -0.5
    bias = torch.randn(Cout, generator=g, device="cuda")
    rbb = torch.randn(B, Cout, generator=g, device="cuda")
    res = _rand_bf16(g, B * H * H, Cout)
    wq = w.to(torch.bfloat16).float()
    pre = F.conv2d(x.float(), wq, bias, padding=1) + rbb[:, :, None, None]
    pre = pre.permute(0, 2, 3, 1).reshape(B * H * H, Cout)
    want = pre + res.float()
    wp = ops.pack_conv_weight(w)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = torch.zeros(B * H * H, Cout, dtype=torch.bfloat16, device="cuda")
    cat = torch.zeros(B * H * H, Cout + 64, dtype=torch.bfloat16, device="cuda")
    cap_pre = torch.zeros(B * H * H, Cout, dtype=torch.float16, device="cuda")
    cap = torch.zeros(B * H * H, Cout, dtype=torch.float16, device="cuda")
    ep = ops.make_epilogue(out=out, bias=bias, row_batch_bias=rbb, rows_per_batch=H * H, residual=res, out2=cat[:, 64:],
                           cap_pre=cap_pre, caps=[(cap, 0, Cout)])
    ops.conv3x3(x_nhwc, wp, ep)
    torch.cuda.synchronize()
    what = "resnet conv B%d %dx%d %d->%d" % (B, H, H, Cin, Cout)
    check_close(cap_pre, pre, what=what + " increment capture")
    check_close(out, want, what=what + " out")
    check_close(cap, want, what=what + " out capture")
    check_close(cat[:, 64:], want, what=what + " concat slice")
    assert float(cat[:, :64].abs().max()) == 0.0
    for name, t in (("out", out), ("out capture", cap)):      # sparse corruption hides in an L2 norm
        worst = float((t.float() - want).abs().max() / want.abs().max())
        assert worst < 3e-2, "%s %s: max-relative error %.3e" % (what, name, worst)


def test_conv3x3_residual_tma_refill_race(cuda_dev):
    """Round-1 defect, root-caused in round 2: the epilogue handed its residual staging buffer back to TMA while the
    shared-memory loads of the current round were still queued behind outstanding global loads, so 16-byte pieces of a
    round's residual were read as the NEXT round's (zeros where that box lies beyond N). About every second launch of
    this shape (conv 64x64, 640 -> 640, residual, B = 2) showed 30-350 wrong elements; 12 launches, per-element check."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, C = 2, 64, 640
    x = _rand_bf16(g, B, C, H, H)
    w = torch.randn(C, C, 3, 3, generator=g, device="cuda") * (9 * C) ** -0.5
    bias = torch.randn(C, generator=g, device="cuda")
    res = _rand_bf16(g, B * H * H, C)
    want = F.conv2d(x.float(), w.to(torch.bfloat16).float(), bias, padding=1).permute(0, 2, 3, 1).reshape(B * H * H, C)
    want = want + res.float()
    wp = ops.pack_conv_weight(w)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    tol = 0.02 * float(want.abs().max())
    for rep in range(12):
        out = torch.zeros(B * H * H, C, dtype=torch.bfloat16, device="cuda")
        ops.conv3x3(x_nhwc, wp, ops.make_epilogue(out=out, bias=bias, residual=res))
        torch.cuda.synchronize()
        bad = int(((out.float() - want).abs() > tol).sum())
        assert bad == 0, "launch %d: %d elements off by more than 2 %% of the largest value" % (rep, bad)


@pytest.mark.parametrize("B,HW,C,silu", [(2, 64 * 64, 320, True), (3, 32 * 32, 1920, True), (2, 128 * 128, 128, True),
                                         (1, 16 * 16, 2560, False), (2, 1024, 960, True), (2, 77, 640, True)])
def test_groupnorm(cuda_dev, B, HW, C, silu):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    x = (torch.randn(B, HW, C, generator=g, device="cuda") * 2 + 0.5).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device="cuda")
    beta = 0.1 * torch.randn(C, generator=g, device="cuda")
    y = ops.groupnorm(x, gamma, beta, 32, 1e-5, silu)
    torch.cuda.synchronize()
    want = F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5)
    if silu:
        want = F.silu(want)
    check_close(y, want.permute(0, 2, 1), what="groupnorm C%d" % C)


@pytest.mark.parametrize("M,C,mod", [(1000, 640, False), (333, 1280, False), (512, 1152, True), (64, 320, False),
                                     (20011, 1280, False), (9002, 1152, True), (70001, 320, False), (300, 3072, True)])
def test_layernorm(cuda_dev, M, C, mod):
    """The last-but-one three exceed the warp count of the persistent kernel (148 x 16): every warp walks several rows with
    the next row prefetched; 3072 columns stay on the single-wave kernel."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(7)
    x = (torch.randn(M, C, generator=g, device="cuda") * 3 + 1).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, generator=g, device="cuda")
    beta = 0.1 * torch.randn(C, generator=g, device="cuda")
    if mod:
        rows = M // 2
        sc = 0.1 * torch.randn(2, C, generator=g, device="cuda")
        sh = 0.1 * torch.randn(2, C, generator=g, device="cuda")
        y = ops.layernorm(x, None, None, 1e-6, sc, sh, rows)
        want = F.layer_norm(x.float(), (C,), None, None, 1e-6)
        want = want * (1 + sc.repeat_interleave(rows, 0)) + sh.repeat_interleave(rows, 0)
    else:
        y = ops.layernorm(x, gamma, beta, 1e-5)
        want = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    check_close(y, want, what="layernorm")


@pytest.mark.parametrize("B,heads,Nq,Nk", [(2, 4, 1024, 1024), (1, 10, 4096, 4096), (2, 5, 1024, 77), (1, 2, 200, 300),
                                           (3, 1, 64, 64), (4, 10, 1024, 1024), (8, 20, 1024, 77), (2, 3, 576, 576)])
@pytest.mark.parametrize("v_f16", [False, True])
def test_attention64(cuda_dev, B, heads, Nq, Nk, v_f16):
    """Persistent tcgen05 / TMEM kernel: v_f16=False rounds P to bf16 (bf16 V, the text cross-attention path),
    v_f16=True keeps P and V in fp16 (self-attention). (4, 10, 1024, 1024) = 160 work items and (8, 20, 1024, 77) = 640
    exceed the SM count, so CTAs walk several items through one continuous tile pipeline."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(8)
    C = heads * 64
    # q/k/v as column slices of a fused projection output (exercises the row pitch)
    qkv = _rand_bf16(g, B * Nq, 3 * C) if Nq == Nk else None
    if qkv is not None:
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q = _rand_bf16(g, B * Nq, C)
        kv = _rand_bf16(g, B * Nk, 2 * C)
        k, v = kv[:, :C], kv[:, C:]
    if v_f16:
        v = v.float().half().contiguous()
    o = ops.attention(q, k, v, B, heads, Nq, Nk, 64 ** -0.5, v_f16=v_f16)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Nq, heads, 64).transpose(1, 2)
    kf = k.float().reshape(B, Nk, heads, 64).transpose(1, 2)
    vf = v.float().reshape(B, Nk, heads, 64).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * Nq, C)
    check_close(o, want, what="attention B%d h%d %dx%d f16=%d" % (B, heads, Nq, Nk, v_f16))


@pytest.mark.parametrize("D,heads,Nq,Nk", [(40, 8, 1024, 1024), (80, 8, 256, 256), (160, 8, 200, 77), (72, 16, 512, 300),
                                          (8, 8, 256, 256), (32, 4, 100, 100), (128, 24, 1100, 1100), (72, 16, 4096, 4096),
                                          (128, 3, 300, 77)])
@pytest.mark.parametrize("v_f16", [False, True])
def test_attention_other_head_dims(cuda_dev, D, heads, Nq, Nk, v_f16):
    """SD-1.5 (40/80), PixArt (72) and Flux (128) head dims through the tcgen05 kernel (64-column blocks zero-filled
    beyond the head dim by TMA); 160 / 8 / 32 through the padded-head-dim mma.sync kernel."""
    ops = _ops()
    if v_f16 and D not in (40, 72, 80, 128):
        pytest.skip("fp16 V is a tcgen05-kernel option")
    g = torch.Generator(device="cuda").manual_seed(28)
    B, C = 2, heads * D
    q = _rand_bf16(g, B * Nq, C)
    kv = _rand_bf16(g, B * Nk, 2 * C)
    k, v = kv[:, :C], kv[:, C:]
    if v_f16:
        v = v.float().half().contiguous()
    o = ops.attention(q, k, v, B, heads, Nq, Nk, D ** -0.5, head_dim=D, v_f16=v_f16)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Nq, heads, D).transpose(1, 2)
    kf = k.float().reshape(B, Nk, heads, D).transpose(1, 2)
    vf = v.float().reshape(B, Nk, heads, D).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * Nq, C)
    check_close(o, want, what="attention d=%d" % D)


@pytest.mark.parametrize("D,heads,Nq,Nk", [(72, 16, 512, 300), (72, 8, 64, 24), (64, 4, 200, 77)])
def test_attention_key_bias(cuda_dev, D, heads, Nq, Nk):
    """PixArt masked cross-attention: additive (1 - mask) * -10000 key bias (ragged valid length per batch)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(29)
    B, C = 2, heads * D
    q = _rand_bf16(g, B * Nq, C)
    kv = _rand_bf16(g, B * Nk, 2 * C)
    k, v = kv[:, :C], kv[:, C:]
    mask = torch.ones(B, Nk, device="cuda")
    mask[0, Nk // 3:] = 0
    mask[1, Nk - 1:] = 0
    bias = (1 - mask) * -10000.0
    o = ops.attention_bias(q, k, v, B, heads, Nq, Nk, D ** -0.5, D, bias)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Nq, heads, D).transpose(1, 2)
    kf = k.float().reshape(B, Nk, heads, D).transpose(1, 2)
    vf = v.float().reshape(B, Nk, heads, D).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=bias[:, None, None, :])
    check_close(o, want.transpose(1, 2).reshape(B * Nq, C), what="attention key bias d=%d" % D)


def test_attention_peaked_softmax(cuda_dev):
    """Large score range (peaked rows) through the half2-exp path of the tcgen05 kernel."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(18)
    B, heads, N = 1, 2, 512
    C = heads * 64
    q = _rand_bf16(g, B * N, C, scale=4.0)
    k = _rand_bf16(g, B * N, C, scale=4.0)
    v = torch.randn(B * N, C, generator=g, device="cuda").half()
    o = ops.attention(q, k, v, B, heads, N, N, 64 ** -0.5, v_f16=True)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, N, heads, 64).transpose(1, 2)
    kf = k.float().reshape(B, N, heads, 64).transpose(1, 2)
    vf = v.float().reshape(B, N, heads, 64).transpose(1, 2)
    want = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * N, C)
    check_close(o, want, tol=2e-2, what="peaked attention")


def test_linear_mixed_output_dtype(cuda_dev):
    """Fused QKV projection whose V columns are written as fp16 bit patterns (out_f16_from)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(19)
    M, C = 700, 128
    a = _rand_bf16(g, M, C)
    w = _rand_bf16(g, 3 * C, C, scale=C ** -0.5)
    out = torch.zeros(M, 3 * C, dtype=torch.bfloat16, device="cuda")
    ops.linear(a, w, ops.make_epilogue(out=out, out_f16_from=2 * C))
    torch.cuda.synchronize()
    want = a.float() @ w.float().T
    check_close(out[:, :2 * C], want[:, :2 * C], what="q|k columns (bf16)")
    v16 = out[:, 2 * C:].contiguous().view(torch.float16)
    check_close(v16, want[:, 2 * C:], what="v columns (fp16)")


def test_softmax_rows(cuda_dev):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(9)
    s = (torch.randn(300, 2048, generator=g, device="cuda") * 4).to(torch.bfloat16)
    want = torch.softmax(s.float(), dim=-1)
    ops.softmax_rows_(s)
    torch.cuda.synchronize()
    check_close(s, want, what="softmax rows")


def test_upsample_and_im2col(cuda_dev):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(10)
    x = _rand_bf16(g, 2, 16, 16, 64)
    y = ops.upsample_nearest2x(x)
    torch.cuda.synchronize()
    want = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), want), "nearest upsample must be exact"
    img = torch.rand(2, 3, 32, 32, generator=g, device="cuda") * 2 - 1
    a = ops.im2col_small(img, nchw_f32=True)
    torch.cuda.synchronize()
    cols = F.unfold(img, 3, padding=1)  # (B, C*9, L) with index c*9 + tap
    cols = cols.view(2, 3, 9, 32 * 32).permute(0, 3, 2, 1).reshape(2 * 32 * 32, 27)  # -> tap*3 + c
    assert torch.equal(a[:, :27].float(), cols.to(torch.bfloat16).float())
    assert a[:, 27:].abs().max().item() == 0


@pytest.mark.parametrize("out_hw", [(128, 128), (48, 48)])
def test_resize_concat(cuda_dev, out_hw):
    """aggregation_network.py:62-66: F.interpolate(f, size, mode='bilinear') for every map, then cat."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(11)
    B = 2
    maps_nchw = [torch.randn(B, 1280, 32, 32, generator=g, device="cuda").half(),
                 torch.randn(B, 640, 64, 64, generator=g, device="cuda").half(),
                 torch.randn(B, 64, 128, 128, generator=g, device="cuda").half()]
    maps = [m.permute(0, 2, 3, 1).reshape(B, -1, m.shape[1]).contiguous() for m in maps_nchw]
    r = ops.resize_concat(maps, out_hw, nhwc=True, nchw=True, with_sumsq=True)
    torch.cuda.synchronize()
    want = torch.cat([F.interpolate(m, out_hw, mode="bilinear") for m in maps_nchw], dim=1)  # fp16 like the reference
    got_nchw = r["nchw"]
    got_nhwc = r["nhwc"].view(B, out_hw[0], out_hw[1], -1).permute(0, 3, 1, 2)
    # fp16 outputs of the same fp32 interpolation: allow 1 ulp of fp16
    assert (got_nchw.float() - want.float()).abs().max().item() <= 4e-3
    assert torch.equal(got_nchw, got_nhwc.contiguous()), "NHWC and NCHW stacks must hold identical values"
    ss = (r["nhwc"].float() ** 2).sum(-1)
    assert rel_err(r["sumsq"], ss) < 1e-5


def _oracle_corr(f1, f2, pts, load):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import O
    return O.find_nn_source_correspondences(f1.float().cpu(), f2.float().cpu(), pts, load)


def _check_corr(points2, f1, f2, pts, load):
    want, sims = _oracle_corr(f1, f2, pts, load)
    got = points2.cpu()
    agree = (got == want).all(dim=-1)
    # a disagreement is only acceptable as a documented near-tie: the oracle's similarity at our position is within
    # fp16 resolution (2^-10 relative) of its maximum
    flat = got[:, 0] * load[0] + got[:, 1]
    ours = sims[torch.arange(len(flat)), flat]
    near_tie = (sims.max(dim=-1).values - ours) <= 2 ** -10
    assert bool((agree | near_tie).all()), "arg-max differs beyond a near-tie"
    assert agree.float().mean().item() >= 0.995, "agreement %.4f < 99.5%%" % agree.float().mean().item()


def test_correspondence_golden(cuda_dev):
    """vs tests/golden/correspondence.pt, produced by the reference's correspondence_utils (tools/make_golden.py)."""
    import os
    from generic_diffusion_feature_b200 import correspondence as C
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "correspondence.pt"),
                      weights_only=False)
    pts = gold["points"].numpy()
    load = tuple(gold["load_size"])
    _, p2 = C.find_nn_source_correspondences(gold["f1"].cuda().half(), gold["f2"].cuda().half(), pts, None, load)
    torch.cuda.synchronize()
    agree = (p2.cpu() == gold["points2"]).all(dim=-1).float().mean().item()
    assert agree >= 0.95, agree          # fixture inputs are fp32; ours are fp16-rounded copies
    _check_corr(p2, gold["f1"].half(), gold["f2"].half(), pts, load)
    assert np.array_equal(C.points_to_idxs(pts, load), gold["idx"].numpy())


@pytest.mark.parametrize("C,hw,load,n", [(256, 32, 128, 300), (3840, 128, 512, 512)])
def test_correspondence_vs_oracle(cuda_dev, C, hw, load, n):
    from generic_diffusion_feature_b200 import correspondence as Cm
    g = torch.Generator().manual_seed(21)
    # smooth-ish random stacks so that neighbouring positions are correlated, like real feature maps
    base = torch.randn(1, C, hw // 4, hw // 4, generator=g)
    f1 = (F.interpolate(base, (hw, hw), mode="bilinear") + 0.3 * torch.randn(1, C, hw, hw, generator=g)).half()
    f2 = (F.interpolate(base, (hw, hw), mode="bilinear") + 0.3 * torch.randn(1, C, hw, hw, generator=g)).half()
    pts = np.random.RandomState(3).uniform(0, load - 1, size=(n, 2))
    _, p2 = Cm.find_nn_source_correspondences(f1.cuda(), f2.cuda(), pts, None, (load, load))
    torch.cuda.synchronize()
    _check_corr(p2, f1, f2, pts, (load, load))


@pytest.mark.parametrize("C,G,HW,B", [(128, 32, 4096, 2), (256, 32, 1024, 3), (512, 32, 256, 2)])
def test_conv_fused_groupnorm_stats(cuda_dev, C, G, HW, B):
    """GroupNorm statistics accumulated by the producing conv's epilogue (gn_sums) == statistics of its output."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(41)
    H = W = int(math.isqrt(HW))
    x = _rand_bf16(g, B, H, W, 64)
    w = torch.randn(C, 64, 3, 3, generator=g, device="cuda") * (9 * 64) ** -0.5
    bias = torch.randn(C, generator=g, device="cuda")
    res = _rand_bf16(g, B * HW, C)
    wp = ops.pack_conv_weight(w)
    out = torch.zeros(B * HW, C, dtype=torch.bfloat16, device="cuda")
    sums = torch.zeros(B, G, 2, device="cuda")
    ep = ops.make_epilogue(out=out, bias=bias, residual=res, gn_sums=sums, gn_cpg=C // G, gn_groups=G, gn_rows_per_img=HW)
    ops.conv3x3(x, wp, ep)
    torch.cuda.synchronize()
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), bias, padding=1)
    want = want.permute(0, 2, 3, 1).reshape(B * HW, C) + res.float()
    check_close(out, want, what="conv with fused GN stats")
    wg = want.reshape(B, HW, G, C // G)
    s_want = wg.sum(dim=(1, 3))
    q_want = (wg * wg).sum(dim=(1, 3))
    assert torch.allclose(sums[..., 0], s_want, rtol=2e-3, atol=0.5), (sums[..., 0] - s_want).abs().max()
    assert torch.allclose(sums[..., 1], q_want, rtol=2e-3, atol=0.5), (sums[..., 1] - q_want).abs().max()


@pytest.mark.parametrize("ratio", [2, 3])
def test_feature_resize_pool_vs_reference_store(cuda_dev, ratio):
    """gdf_op_avgpool_nhwc (through pool_views) vs the fixture written by the reference's real FeatureStore.store with
    resize_ratio 2 / 3 (feature_extractor.py:51-53); inputs rounded to fp16 first, as the arena holds them."""
    import os
    from common import ROOT
    from generic_diffusion_feature_b200 import _lib
    from generic_diffusion_feature_b200.components.feature_extractor import pool_views
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "feature_store_resize.pt"), weights_only=False)
    conv = gold["conv"].half().cuda()                                   # (B, C, h, w)
    vit = gold["vit"].half().cuda()                                     # (B, N, C) token-major
    feats = {"a": conv.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2),
             "b": vit.view(2, 16, 16, 128).permute(0, 3, 1, 2)}
    got = pool_views(_lib.load(), feats, ratio)
    torch.cuda.synchronize()
    assert list(got.keys()) == ["a", "b"]
    # a 4-D attention map (B, heads, Nq, Nk) is stored NCHW-contiguous and pooled over its last two axes
    amap = torch.rand(2, 4, 16, 77, device="cuda").half()
    pm = pool_views(_lib.load(), {"m-self-map": amap}, ratio)["m-self-map"]
    want_m = F.adaptive_avg_pool2d(amap.float(), (16 // ratio, 77 // ratio))
    assert pm.shape == want_m.shape and (pm.float() - want_m).abs().max().item() < 2e-3
    for k in ("a", "b"):
        want = gold["r%d" % ratio][k]
        assert got[k].shape == want.shape and got[k].dtype == torch.float16
        assert (got[k].float().cpu() - want).abs().max().item() < 4e-3


@pytest.mark.gpu
def test_aggregation_head_conv_on_fp16_stack(cuda_dev):
    """AggregationNetwork.out (3x3, bias-free, aggregation_network.py:22,97-99) on the fp16 stack
    through the implicit-GEMM conv with fp16 operands vs F.conv2d in fp32 on the same stack values and weights rounded
    to fp16 (tight), and vs unrounded fp32 weights (fp16 weight rounding only)."""
    from generic_diffusion_feature_b200 import correspondence as C
    g = torch.Generator(device="cuda").manual_seed(31)
    B, H, dim, out_dim = 2, 32, 384, 192
    stack = (torch.randn(B, H * H, dim, generator=g, device="cuda") * 0.5).half()
    w = torch.randn(out_dim, dim, 3, 3, generator=g, device="cuda") / (9 * dim) ** 0.5
    head = C.AggregationHead(w)
    got = head(stack, (H, H))
    torch.cuda.synchronize()
    x = stack.float().view(B, H, H, dim).permute(0, 3, 1, 2)
    want16 = F.conv2d(x, w.half().float(), padding=1)
    want32 = F.conv2d(x, w, padding=1)
    assert got.shape == want32.shape
    assert rel_err(got, want16) < 2e-4
    assert rel_err(got, want32) < 2e-3


@pytest.mark.gpu
def test_segmentor_feature_head_matches_reference(cuda_dev):
    """DiffusionSegmentor.extract_feat's ResBlocks (segmentation/models/diffusion_segmentor.py:23-53,
    232-246, eval mode) on fp16 maps through the fp16-operand tcgen05 convolution (ReLU epilogue, fp16 residual, channel
    slices of the concat buffer) vs the outputs of the reference's REAL module (tests/golden/segmentor_head.pt, fp32
    cuDNN-free CPU run) and vs the CPU oracle. fp16 operands + fp16 intermediates: tolerance 3e-3 of the map's range."""
    from common import O
    from generic_diffusion_feature_b200 import segmentation as S
    gold = torch.load(os.path.join(GOLD, "segmentor_head.pt"), weights_only=False)
    sd = {k: v.float() for k, v in gold["state_dict"].items()}
    feats = {k: v.cuda() for k, v in gold["features"].items()}
    head = S.SegmentorFeatureHead(gold["feature_layers"], sd)
    outs = head(feats)
    torch.cuda.synchronize()
    o_outs = O.seg_extract_feat(gold["features"], gold["feature_layers"], sd)
    assert len(outs) == len(gold["outs"])
    for got, want, o in zip(outs, gold["outs"], o_outs):
        assert got.shape == want.shape and got.dtype == torch.float32
        assert rel_err(got.cpu(), want.float()) < 3e-3
        assert rel_err(got.cpu(), o) < 3e-3
    # MultiRes: ONE ResBlock applied n times (diffusion_segmentor.py:46-53)
    pre = "up_level1_upsampler_out."
    mr = S.MultiRes({"m.res.0." + k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, "m", gold["multires_n"])
    x = feats["up-level1-upsampler-out"].permute(0, 2, 3, 1).contiguous()
    y = mr(x).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    assert rel_err(y.cpu(), gold["multires_out"].float()) < 5e-3
    with pytest.raises(ValueError):
        S.ResBlock({k.replace(pre, "q."): v[:32, :32] if v.dim() == 4 else v[:32] for k, v in sd.items()
                    if k.startswith(pre)}, "q")
    # several-extractors branch (diffusion_segmentor.py:248-297) vs the reference's real extract_feat: up to 4 + 2 + 1
    # ResBlocks in sequence with fp16 intermediates
    mg = gold["multi"]
    msd = {k: v.float() for k, v in mg["state_dict"].items()}
    mhead = S.MultiSegmentorFeatureHead(mg["feature_layers"], mg["c_per_level"], msd)
    mouts = mhead([{k: v.cuda() for k, v in f.items()} for f in mg["features"]])
    torch.cuda.synchronize()
    assert len(mouts) == len(mg["outs"])
    for got, want in zip(mouts, mg["outs"]):
        assert got.shape == want.shape
        assert rel_err(got.cpu(), want.float()) < 8e-3


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,W,N,G", [(2, 128, 128, 128, 32), (1, 64, 256, 64, 0), (3, 96, 384, 128, 8)])
def test_conv_in_fused(cuda_dev, B, H, W, N, G):
    """VAE Encoder.conv_in fused from the fp32 NCHW image (the 27-tap operand is built in shared memory): vs F.conv2d on
    bf16-rounded image / weights, and the GroupNorm sums it accumulates for the following GroupNorm."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(41)
    img = torch.rand(B, 3, H, W, generator=g, device="cuda") * 2 - 1
    w = torch.randn(N, 3, 3, 3, generator=g, device="cuda") / 27 ** 0.5
    bias = torch.randn(N, generator=g, device="cuda") * 0.1
    out, sums = ops.conv_in_fused(img, w, bias, gn_groups=G)
    torch.cuda.synchronize()
    want = F.conv2d(img.bfloat16().float(), w.bfloat16().float(), bias, padding=1)       # (B, N, H, W) fp32
    got = out.float().view(B, H, W, N).permute(0, 3, 1, 2)
    check_close(got, want, what="conv_in fused")
    if G:
        wg = want.view(B, G, N // G, H * W)
        ref = torch.stack([wg.sum(dim=(2, 3)), (wg * wg).sum(dim=(2, 3))], dim=-1)
        assert rel_err(sums, ref) < 1e-3
