"""GPU: kernel-level checks of the per-image ConvBlock tails' hand-over pieces (csrc/convblock_fused.cu,
csrc/pointwise.cu), called through the C ABI:

  * pcm_maxpool2_bwd_skip_dot — the backward of MaxPool2d(2) (+ time-mean skip; reference
    src/unet_convlstm_attention.py:21-24,91-93, src/unet.py:54) that also emits the per-pixel sum
    sdot = sum_c dx*x for the ConvBlock backward that follows — against torch autograd on the CPU, on even and ODD
    grids (floor pooling: the last row / column is not pooled but still receives the skip gradient);
  * pcm_convblock_tail_bwd_sdot with that sum against the same kernel streaming dout and out itself (sdot = NULL),
    with the shared-memory scratch on and off: every output of the backward tail must agree.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import pcm_b200  # noqa: F401
    from pcm_b200 import ops
    return ops


POOL_SHAPES = [  # (B, T, H, W, C, with_dy, with_skip)
    (2, 3, 8, 12, 16, True, True), (2, 3, 7, 9, 16, True, True), (1, 2, 5, 6, 32, True, False), (3, 1, 6, 7, 64, True, True),
    (2, 2, 9, 4, 8, False, True), (1, 1, 3, 3, 256, True, True), (4, 6, 48, 72, 16, True, True),
]


@pytest.mark.parametrize("shape", POOL_SHAPES)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_maxpool2_bwd_skip_dot_matches_autograd(ops, shape, dt):
    B, T, H, W, C, with_dy, with_skip = shape
    N = B * T
    g = torch.Generator().manual_seed(H * 100 + W * 10 + C + T)
    # values on a coarse grid: every 2x2 window has ties, so "first maximum wins" is exercised
    x = (torch.randint(-3, 4, (N, H, W, C), generator=g).float() / 4).to(dt)
    dy = torch.randn(N, H // 2, W // 2, C, generator=g).to(dt) if with_dy else None
    dskip = torch.randn(B, H, W, C, generator=g).to(dt) if with_skip else None
    # reference: t-major frames (image n = t*B + b); pooled = maxpool(x), skip = mean over t
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    loss = 0.0
    if with_dy:
        loss = loss + (F.max_pool2d(xr, 2) * dy.float().permute(0, 3, 1, 2)).sum()
    if with_skip:
        skip = xr.reshape(T, B, C, H, W).mean(0)
        loss = loss + (skip * dskip.float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    want_dx = xr.grad.permute(0, 2, 3, 1).contiguous()
    xg = x.cuda()
    dyg = dy.cuda() if with_dy else None
    dsg = dskip.cuda() if with_skip else None
    d = ops._DT[dt]
    outs = []
    for use_dot in (False, True):
        dx = torch.empty_like(xg)
        sdot = torch.full((N * H * W,), float("nan"), device="cuda") if use_dot else None
        ns, ps = (H * W * C, C) if with_skip else (0, 0)
        ops._call("pcm_maxpool2_bwd_skip_dot", xg.data_ptr(), ops._p(dyg), ops._p(dsg), ns, ps, dx.data_ptr(), ops._p(sdot),
                  N, H, W, C, T, 1, d, ops._s())
        torch.cuda.synchronize()
        outs.append((dx.cpu(), None if sdot is None else sdot.cpu()))
    (dx0, _), (dx1, sdot) = outs
    assert torch.equal(dx0, dx1)                                   # the extra output does not change the gradient
    tol = 1e-6 if dt == torch.float32 else 8e-3
    assert float((dx1.float() - want_dx).abs().max()) <= tol * max(1.0, float(want_dx.abs().max()))
    # sdot is defined on the gradient AS STORED (rounded to dt) times x
    want_s = (dx1.float() * x.float()).sum(-1).reshape(-1)
    assert torch.isfinite(sdot).all()
    assert float((sdot - want_s).abs().max()) <= 1e-5 * max(1.0, float(want_s.abs().max()))


TAIL_SHAPES = [(5, 48, 72, 16), (3, 24, 36, 32), (4, 12, 18, 64), (6, 6, 9, 128), (2, 7, 9, 16), (150, 12, 18, 64)]


@pytest.mark.parametrize("shape", TAIL_SHAPES)
@pytest.mark.parametrize("scratch", ["1", "0"])
def test_tail_bwd_with_supplied_gate_sum_matches_streaming_path(ops, shape, scratch, monkeypatch):
    N, H, W, C = shape
    P, Cr, d = H * W, C // 8, ops._DT[torch.bfloat16]
    monkeypatch.setenv("PCM_TAIL_SCRATCH", scratch)
    g = torch.Generator().manual_seed(N + H + W + C)
    bf = lambda *sh: torch.randn(*sh, generator=g).bfloat16().cuda()
    f32 = lambda *sh: torch.randn(*sh, generator=g).cuda()
    x, dout = bf(N, P, C), bf(N, P, C)
    gamma, beta = 1.0 + 0.1 * f32(C), 0.1 * f32(C)
    w1, w2, wsp = f32(Cr * C) / C ** 0.5, f32(C * Cr) / Cr ** 0.5, f32(98) / 7
    out = torch.empty_like(x)
    stats, pool, se, hid = (torch.zeros(n, device="cuda") for n in (N * 16, N * C, N * C, N * Cr))
    maps, ties = torch.zeros(N * P * 3, device="cuda"), torch.zeros(N * P, device="cuda", dtype=torch.uint8)
    if not ops.lib()._fn["pcm_convblock_fused_supported"](H, W, C, Cr, d):
        pytest.skip("image does not fit one SM")
    ops._call("pcm_convblock_tail_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), w2.data_ptr(),
              wsp.data_ptr(), stats.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(), maps.data_ptr(), ties.data_ptr(),
              out.data_ptr(), N, H, W, C, Cr, 1e-5, d, ops._s())
    sdot = (dout.float() * out.float()).sum(-1).reshape(-1).contiguous()
    res = []
    for s_ptr in (0, sdot.data_ptr()):
        dx = torch.empty_like(x)
        grads = [torch.zeros(n, device="cuda") for n in (C, C, Cr * C, C * Cr, 98)]
        ops._call("pcm_convblock_tail_bwd_sdot", dout.data_ptr(), x.data_ptr(), out.data_ptr(), stats.data_ptr(), gamma.data_ptr(),
                  beta.data_ptr(), w1.data_ptr(), w2.data_ptr(), wsp.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(),
                  maps.data_ptr(), ties.data_ptr(), dx.data_ptr(), *[t.data_ptr() for t in grads], 0, s_ptr, N, H, W, C, Cr,
                  1e-5, d, ops._s())
        torch.cuda.synchronize()
        res.append([dx.float()] + grads)
    for a, b in zip(*res):
        den = float(a.norm()) + 1e-12
        assert float((a - b).norm()) / den < 2e-3, float((a - b).norm()) / den      # bf16 outputs; fp32 sums in another order
