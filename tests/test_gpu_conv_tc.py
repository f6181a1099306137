"""GPU: the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu, called through the C ABI) against the
general-shape SIMT gather kernel and against torch's CPU conv on the same bf16-rounded operands."""
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


def _ref(x, w, bias):
    """x (N,H,W,Ci) bf16-rounded, w (Co,Ci,3,3) bf16-rounded -> NHWC fp32 via torch CPU (fp64 accumulate)."""
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None if bias is None else bias.double(), padding=1)
    return y.permute(0, 2, 3, 1).float()


SHAPES = [  # (N, H, W, Cin, Cout): every 3x3 layer geometry of unet_convlstm_attention (config 3) + ragged cases
    (5, 48, 72, 16, 16), (3, 24, 36, 16, 32), (3, 24, 36, 32, 32), (4, 12, 18, 32, 64), (4, 12, 18, 64, 64),
    (7, 6, 9, 64, 128), (7, 6, 9, 128, 128), (6, 6, 9, 128, 256), (5, 6, 9, 64, 256), (3, 6, 9, 256, 128),
    (3, 6, 9, 256, 64), (2, 12, 18, 128, 64), (2, 24, 36, 64, 32), (2, 48, 72, 32, 16),
    (3, 23, 45, 64, 64), (1, 7, 200, 32, 48), (2, 5, 5, 16, 16), (1, 130, 3, 16, 32),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_conv3x3_tc_matches_reference(ops, shape):
    from pcm_b200._lib import lib
    N, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(N * 1000 + H * 10 + Ci + Co)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).bfloat16()
    bias = torch.randn(Co, generator=g)
    want = _ref(x, w, bias)
    xg, wg, bg = x.cuda(), w.float().cuda(), bias.cuda()
    assert ops.tc_supported(torch.bfloat16, Ci, Co)
    wk = ops.conv_weight_fwd(wg, torch.bfloat16)
    got = ops.conv_s1(xg, wk, N, H, W, Ci, Co, dst_f32=True, bias=bg)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    err = float((got.cpu() - want).norm() / want.norm())
    assert err < 2e-6, err
    # SIMT kernel on the same operands (fp32 accumulate in a different order)
    simt = ops.conv_gather(xg, wk, N, H, W, Ci, H, W, Co, 3, 3, 1, 1, 0, dst_f32=True, bias=bg)
    assert float((got - simt).norm() / simt.norm()) < 2e-6
    # bf16 destination
    got16 = ops.conv_s1(xg, wk, N, H, W, Ci, Co, bias=bg)
    assert float((got16.float().cpu() - want).norm() / want.norm()) < 4e-3


GROUPED = [  # (N, H, W, Cin, Cout, group): the thin layers of config 3 in pixel-group form (pcm_conv3x3_tc_grouped)
    (5, 48, 72, 16, 16, 4), (5, 48, 72, 16, 16, 2), (3, 24, 36, 16, 32, 2), (3, 24, 36, 32, 32, 2), (2, 48, 72, 32, 16, 2),
    (2, 5, 8, 16, 16, 4), (1, 7, 6, 16, 16, 2), (150, 48, 72, 16, 16, 4),
]


@pytest.mark.parametrize("shape", GROUPED)
def test_conv3x3_tc_grouped_matches_reference(ops, shape):
    """Pixel-group form: g adjacent pixels as one GEMM row (g x fewer, g x longer TMA rows).  Forward and the flipped
    data-gradient packing, against fp64 convolution of the same bf16-rounded operands and against the plain form; the
    batched re-pack (PackPlan) must reproduce the grouped packing bit for bit."""
    from pcm_b200._lib import lib
    N, H, W, Ci, Co, grp = shape
    g = torch.Generator().manual_seed(N * 1000 + H * 10 + Ci + Co + grp)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).bfloat16()
    want = _ref(x, w, None)
    xg, wg = x.cuda(), w.float().cuda()
    wk = ops.conv_weight_fwd(wg, torch.bfloat16, group=grp)
    assert tuple(wk.shape) == (9, grp * Co, grp * Ci)
    got = ops.conv_s1(xg, wk, N, H, W, Ci, Co, dst_f32=True, group=grp)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert float((got.cpu() - want).norm() / want.norm()) < 2e-6
    plain = ops.conv_s1(xg, ops.conv_weight_fwd(wg, torch.bfloat16), N, H, W, Ci, Co)
    got16 = ops.conv_s1(xg, wk, N, H, W, Ci, Co, group=grp)
    assert float((got16.float() - plain.float()).norm() / plain.float().norm()) < 1e-3     # bf16 rounding of ~equal fp32 sums
    assert float((got16.float().cpu() - want).norm() / want.norm()) < 4e-3
    # data gradient of the same layer through the flipped grouped packing
    dy = torch.randn(N, H, W, Co, generator=g).bfloat16()
    xz = torch.zeros(N, Ci, H, W, requires_grad=True, dtype=torch.float64)
    F.conv2d(xz, w.double(), padding=1).backward(dy.double().permute(0, 3, 1, 2))
    wkt = ops.conv_weight_dgrad(wg, torch.bfloat16, group=grp)
    dx = ops.conv_s1(dy.cuda(), wkt, N, H, W, Co, Ci, dst_f32=True, group=grp)
    wantdx = xz.grad.permute(0, 2, 3, 1).float()
    assert float((dx.cpu() - wantdx).norm() / wantdx.norm()) < 2e-6
    # the one-launch re-pack of the trainer writes the same grouped kernel
    plan = ops.PackPlan()
    with ops.use_pack_plan(plan):
        a = ops.conv_weight_fwd(wg, torch.bfloat16, group=grp)
        b = ops.conv_weight_dgrad(wg, torch.bfloat16, group=grp)
        a.zero_(); b.zero_()
        plan.repack()
    torch.cuda.synchronize()
    assert torch.equal(a, wk) and torch.equal(b, wkt)


def test_conv3x3_tc_accumulate_and_views(ops):
    """fp32 accumulate (ConvLSTM Wx.x + Wh.h split), time-strided source images, channel-slice source view."""
    T, B, H, W, Ci, Co = 3, 4, 6, 9, 64, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, H, W, Ci, generator=g).bfloat16()          # image n = b*T + t
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / 24).bfloat16()
    base = torch.randn(T, B, H, W, Co, generator=g)
    xg, wk = x.cuda(), ops.conv_weight_fwd(w.float().cuda(), torch.bfloat16)
    out = base.clone().cuda()
    img = H * W * Ci
    for t in range(T):
        ops.conv_s1(xg, wk, B, H, W, Ci, Co, dst=out[t], dst_f32=True, accumulate=True, src_ns=T * img, src_off=t * img)
    torch.cuda.synchronize()
    for t in range(T):
        want = base[t] + _ref(x[:, t], w, None)
        assert float((out[t].cpu() - want).norm() / want.norm()) < 2e-6
    # source = channels [32, 64) of a 96-channel concat buffer; destination = channels [16, 48) of a 64-channel one
    cat = torch.randn(2, 12, 18, 96, generator=g).bfloat16()
    w2 = (torch.randn(32, 32, 3, 3, generator=g) / 17).bfloat16()
    wk2 = ops.conv_weight_fwd(w2.float().cuda(), torch.bfloat16)
    dst = torch.zeros(2, 12, 18, 64, dtype=torch.bfloat16).cuda()
    ops.conv_s1(cat.cuda(), wk2, 2, 12, 18, 32, 32, dst=dst, src_ps=96, src_off=32, dst_ps=64, dst_off=16)
    torch.cuda.synchronize()
    want = _ref(cat[..., 32:64], w2, None)
    got = dst.float().cpu()
    assert float((got[..., 16:48] - want).norm() / want.norm()) < 4e-3
    assert float(got[..., :16].abs().max()) == 0.0 and float(got[..., 48:].abs().max()) == 0.0


def test_dgrad_weights_flip(ops):
    """dx = conv(dy, flip(W)^T): the data gradient through the same kernel, vs torch autograd."""
    N, H, W, Ci, Co = 2, 12, 18, 32, 64
    g = torch.Generator().manual_seed(9)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / 17).bfloat16().float()
    dy = torch.randn(N, H, W, Co, generator=g).bfloat16()
    x = torch.zeros(N, Ci, H, W, requires_grad=True, dtype=torch.float64)
    F.conv2d(x, w.double(), padding=1).backward(dy.double().permute(0, 3, 1, 2))
    want = x.grad.permute(0, 2, 3, 1).float()
    wkt = ops.conv_weight_dgrad(w.cuda(), torch.bfloat16)
    for use in ("tc", "simt"):
        if use == "tc":
            got = ops.conv_s1(dy.cuda(), wkt, N, H, W, Co, Ci, dst_f32=True)
        else:
            got = ops.conv_gather(dy.cuda(), wkt, N, H, W, Co, H, W, Ci, 3, 3, 1, 1, 0, dst_f32=True)
        assert float((got.cpu() - want).norm() / want.norm()) < 2e-6, use


WG_SHAPES = [  # (N, H, W, Cin, Cout)
    (5, 48, 72, 16, 16), (3, 24, 36, 16, 32), (3, 24, 36, 32, 32), (4, 12, 18, 32, 64), (4, 12, 18, 64, 64),
    (7, 6, 9, 64, 128), (7, 6, 9, 128, 128), (6, 6, 9, 128, 256), (5, 6, 9, 64, 256), (2, 12, 18, 128, 64),
    (2, 24, 36, 64, 32), (2, 48, 72, 32, 16), (3, 23, 45, 64, 64), (1, 7, 200, 32, 16), (2, 5, 5, 16, 16),
    (300, 6, 9, 128, 256), (40, 48, 72, 16, 16),
]


@pytest.mark.parametrize("shape", WG_SHAPES)
def test_wgrad3x3_tc_matches_reference(ops, shape):
    from pcm_b200._lib import lib
    N, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(N * 7 + H + Ci * 3 + Co)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    dy = (torch.randn(N, H, W, Co, generator=g) / (N * H * W) ** 0.5).bfloat16()
    w = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double().permute(0, 3, 1, 2), w, padding=1).backward(dy.double().permute(0, 3, 1, 2))
    want = w.grad.float()
    assert ops.wgrad_tc_supported(torch.bfloat16, Co, Ci, H, W)
    dw = torch.zeros(Co, Ci, 3, 3, device="cuda")
    ops.conv3x3_wgrad(dy.cuda(), x.cuda(), dw, N, H, W, Co, Ci, Ci)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    err = float((dw.cpu() - want).norm() / want.norm())
    assert err < 1e-5, err
    # accumulates into what is already there; a second call doubles the result
    ops.conv3x3_wgrad(dy.cuda(), x.cuda(), dw, N, H, W, Co, Ci, Ci)
    assert float((dw.cpu() - 2 * want).norm() / want.norm()) < 2e-5


@pytest.mark.parametrize("shape", [(5, 48, 72, 16, 16), (3, 24, 36, 16, 32), (3, 24, 36, 32, 32), (2, 48, 72, 32, 16),
                                   (2, 6, 8, 16, 16), (150, 48, 72, 16, 16)])
@pytest.mark.parametrize("grp", ["0", "2", "4"])
def test_wgrad3x3_tc_grouped_forms_agree(ops, shape, grp, monkeypatch):
    """Pixel-group form of the thin-layer weight gradient (pcm_wgrad3x3_tc_grouped) against fp64 autograd, for every
    group factor, into the parameter layout (with padded input channels: Ci_real < Ci) and into the packed layout."""
    from pcm_b200._lib import lib
    monkeypatch.setenv("PCM_WGRAD_GROUP", grp)
    N, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(N * 7 + H + Ci * 3 + Co)
    Cr = Ci - 9 if Ci == 16 else Ci                      # the first layer has 7 real input channels of 16
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    x[..., Cr:] = 0
    dy = (torch.randn(N, H, W, Co, generator=g) / (N * H * W) ** 0.5).bfloat16()
    w = torch.zeros(Co, Cr, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x[..., :Cr].double().permute(0, 3, 1, 2), w, padding=1).backward(dy.double().permute(0, 3, 1, 2))
    want = w.grad.float()
    dw = torch.zeros(Co, Cr, 3, 3, device="cuda")
    ops.conv3x3_wgrad(dy.cuda(), x.cuda(), dw, N, H, W, Co, Ci, Cr)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert float((dw.cpu() - want).norm() / want.norm()) < 1e-5
    if grp != "0":
        assert ops.wgrad_group(torch.bfloat16, Co, Ci, W) > 1
        packed = torch.zeros(9, Co, Ci, device="cuda")
        dyg, xg = dy.cuda(), x.cuda()
        ops._call("pcm_wgrad3x3_tc_grouped", dyg.data_ptr(), H * W * Co, Co, xg.data_ptr(), H * W * Ci, Ci, Cr,
                  packed.data_ptr(), Ci, 1, Co * Ci, N, H, W, ops.wgrad_group(torch.bfloat16, Co, Ci, W), ops._s())
        torch.cuda.synchronize()
        got = packed.reshape(3, 3, Co, Ci).permute(2, 3, 0, 1)[:, :Cr]
        assert float((got.cpu() - want).norm() / want.norm()) < 1e-5


def test_wgrad3x3_tc_slices(ops):
    """ConvLSTM use: dW[:, ci_off:ci_off+Ci] slice of a (Co, Ci_tot, 3, 3) gradient, time-strided x frames,
    padded input channels (Ci_real < Ci)."""
    T, B, H, W, Cx, Ch = 3, 4, 6, 9, 128, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, H, W, Cx, generator=g).bfloat16()            # image n = b*T + t
    dg = (torch.randn(T, B, H, W, 4 * Ch, generator=g) / 30).bfloat16()
    dw = torch.zeros(4 * Ch, Cx + Ch, 3, 3, device="cuda")
    img = H * W * Cx
    for t in range(T):
        ops.conv3x3_wgrad(dg[t].cuda(), x.cuda(), dw, B, H, W, 4 * Ch, Cx, Cx, Ci_tot=Cx + Ch, x_ns=T * img, x_off=t * img)
    w = torch.zeros(4 * Ch, Cx, 3, 3, dtype=torch.float64, requires_grad=True)
    xs = x.double().permute(1, 0, 4, 2, 3).reshape(T * B, Cx, H, W)
    F.conv2d(xs, w, padding=1).backward(dg.double().reshape(T * B, H, W, 4 * Ch).permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    got = dw.cpu()
    assert float((got[:, :Cx] - w.grad.float()).norm() / w.grad.norm()) < 1e-5
    assert float(got[:, Cx:].abs().max()) == 0.0
    # first layer: 7 real input channels padded to 16
    x7 = torch.zeros(3, 12, 18, 16)
    x7[..., :7] = torch.randn(3, 12, 18, 7, generator=g)
    x7 = x7.bfloat16()
    dy = (torch.randn(3, 12, 18, 16, generator=g) / 25).bfloat16()
    dw7 = torch.zeros(16, 7, 3, 3, device="cuda")
    ops.conv3x3_wgrad(dy.cuda(), x7.cuda(), dw7, 3, 12, 18, 16, 16, 7)
    w7 = torch.zeros(16, 7, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x7[..., :7].double().permute(0, 3, 1, 2), w7, padding=1).backward(dy.double().permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    assert float((dw7.cpu() - w7.grad.float()).norm() / w7.grad.norm()) < 1e-5


C11_SHAPES = [  # (N, H, W, Cin, Cout): ResidualBlock skip convs, transformer linears (tokens as a 1 x L image)
    (2, 48, 72, 64, 128), (2, 24, 36, 128, 256), (3, 1, 216, 128, 192), (3, 1, 216, 128, 256), (3, 1, 216, 256, 128),
    (5, 1, 216, 128, 128), (2, 7, 11, 32, 64), (1, 1, 1000, 64, 16), (2, 12, 18, 16, 32),
]


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("shape", C11_SHAPES)
def test_conv1x1_tc_matches_reference(ops, shape, relu):
    from pcm_b200 import ops_nn
    from pcm_b200._lib import lib
    N, H, W, Ci, Co = shape
    g = torch.Generator().manual_seed(N * 31 + H + Ci + Co)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    w = (torch.randn(Co, Ci, generator=g) / Ci ** 0.5).bfloat16()
    bias = torch.randn(Co, generator=g)
    want = x.double() @ w.double().t() + bias.double()
    if relu:
        want = want.clamp_min(0)
    wk = w.cuda().reshape(1, Co, Ci).contiguous()
    got = ops_nn.conv_same(x.cuda(), wk, N, H, W, Ci, Co, 1, bias=bias.cuda(), relu=relu)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert lib().last_call == "pcm_conv1x1_tc"
    err = float((got.double().cpu() - want).norm() / want.norm())
    assert err < 4e-3, err                                      # bf16 destination rounding


@pytest.mark.parametrize("shape", C11_SHAPES + [(2, 48, 72, 512, 128), (4, 1, 216, 128, 384)])
def test_wgrad1x1_tc_matches_reference(ops, shape):
    from pcm_b200 import ops_nn
    from pcm_b200._lib import lib
    N, H, W, Ci, Co = shape
    if not (Co in (16, 32, 64) or Co % 128 == 0) or W + 2 > 256:
        pytest.skip("not a tensor-core weight-gradient shape (falls back to the SIMT kernel)")
    g = torch.Generator().manual_seed(N * 17 + W + Ci * 3 + Co)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    dy = (torch.randn(N, H, W, Co, generator=g) / (N * H * W) ** 0.5).bfloat16()
    want = dy.double().reshape(-1, Co).t() @ x.double().reshape(-1, Ci)
    dw = torch.zeros(Co, Ci, 1, 1, device="cuda")
    ops_nn.wgrad_same(dy.cuda(), x.cuda(), dw, N, H, W, Co, Ci, Ci, 1)
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert lib().last_call == "pcm_wgrad1x1_tc"
    err = float((dw.cpu().reshape(Co, Ci).double() - want).norm() / want.norm())
    assert err < 1e-5, err


def test_conv3x3_tc_wide_channels(ops):
    """SimpleCNN widths: Cout 512 (split over two launches), Cin 512 (eight K chunks; weight gradient in 256-wide slices)."""
    from pcm_b200 import ops_nn
    N, H, W, Ci, Co = 1, 12, 18, 512, 512
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, H, W, Ci, generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).bfloat16()
    bias = torch.randn(Co, generator=g)
    want = _ref(x, w, bias)
    wk = ops.conv_weight_fwd(w.float().cuda(), torch.bfloat16)
    got = ops_nn.conv_same(x.cuda(), wk, N, H, W, Ci, Co, 3, bias=bias.cuda())
    assert float((got.float().cpu() - want).norm() / want.norm()) < 4e-3
    dy = (torch.randn(N, H, W, Co, generator=g) / (H * W) ** 0.5).bfloat16()
    wz = torch.zeros(Co, Ci, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double().permute(0, 3, 1, 2), wz, padding=1).backward(dy.double().permute(0, 3, 1, 2))
    dw = torch.zeros(Co, Ci, 3, 3, device="cuda")
    ops_nn.wgrad_same(dy.cuda(), x.cuda(), dw, N, H, W, Co, Ci, Ci, 3)
    torch.cuda.synchronize()
    assert float((dw.cpu() - wz.grad.float()).norm() / wz.grad.norm()) < 1e-5


CT_SHAPES = [  # (B, h, w, Ci, Co, Cs): the three Up blocks of config 3, the transformer decoder, ragged sizes
    (4, 6, 9, 64, 64, 64), (3, 12, 18, 64, 32, 32), (2, 24, 36, 32, 16, 16), (2, 12, 18, 128, 64, 0), (2, 24, 36, 64, 32, 0),
    (3, 5, 7, 32, 16, 16), (1, 20, 33, 64, 64, 32),
]


@pytest.mark.parametrize("shape", CT_SHAPES)
def test_convT2x2_tc_matches_reference(ops, shape):
    """ConvTranspose2d(k2,s2) forward (pixel-shuffle epilogue into a concat buffer), data gradient and weight
    gradient (stride-2 TMA views) against torch's CPU conv_transpose2d on the same bf16-rounded operands."""
    from pcm_b200._lib import lib
    B, h, w, Ci, Co, Cs = shape
    g = torch.Generator().manual_seed(B * 13 + h + Ci + Co)
    x = torch.randn(B, h, w, Ci, generator=g).bfloat16()
    wt = (torch.randn(Ci, Co, 2, 2, generator=g) / Ci ** 0.5).bfloat16()
    bias = torch.randn(Co, generator=g)
    Cc = Co + Cs
    H, W = 2 * h, 2 * w
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.double().requires_grad_(True)
    want = F.conv_transpose2d(xr, wr, bias.double(), stride=2)                       # (B, Co, H, W)
    cat = torch.full((B, H, W, Cc), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.convT2x2_fwd(x.cuda(), wt.float().cuda(), bias.cuda(), B, h, w, Ci, Co, cat, H * W * Cc, Cc)
    torch.cuda.synchronize()
    assert lib().last_call == "pcm_convT2x2_tc"
    assert lib()._fn["pcm_tc_error_count"]() == 0
    got = cat[..., :Co].float().cpu()
    assert float((got - want.detach().permute(0, 2, 3, 1).float()).norm() / want.norm()) < 4e-3
    if Cs:
        assert bool((cat[..., Co:] == 7.0).all())                                    # the skip half is untouched
    # gradients from a bf16 dcat whose first Co channels are the convT gradient
    dcat = (torch.randn(B, H, W, Cc, generator=g) / (B * H * W) ** 0.5).bfloat16()
    want.backward(dcat[..., :Co].double().permute(0, 3, 1, 2))
    dx = ops.convT2x2_dgrad(dcat.cuda(), wt.float().cuda(), B, h, w, Ci, Co, H * W * Cc, Cc)
    assert lib().last_call == "pcm_convT2x2_dgrad_tc"
    gwt = torch.zeros(Ci, Co, 2, 2, device="cuda")
    ops.convT2x2_wgrad(x.cuda(), dcat.cuda(), gwt, B, h, w, Ci, Co, H * W * Cc, Cc)
    assert lib().last_call == "pcm_convT2x2_wgrad_tc"
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    wdx = xr.grad.permute(0, 2, 3, 1).float()
    assert float((dx.float().cpu() - wdx).norm() / wdx.norm()) < 4e-3
    assert float((gwt.cpu() - wr.grad.float()).norm() / wr.grad.norm()) < 1e-5


@pytest.mark.parametrize("cfg", [(3, 216, 4), (2, 224, 2), (2, 100, 4), (1, 17, 1)])
def test_mha_fwd_tc_matches_reference(ops, cfg):
    """tcgen05 attention forward (head dim 32) against fp64 softmax(q k^T / sqrt(d)) v on the same bf16 inputs, and the
    saved log-sum-exp against the SIMT kernel's."""
    from pcm_b200._lib import lib
    from pcm_b200.ops import _call, _s
    B, L, nh = cfg
    D, E = 32, 32 * nh
    g = torch.Generator().manual_seed(B * 100 + L + nh)
    qkv = torch.randn(B, L, 3 * E, generator=g).bfloat16()
    q, k, v = [t.double().reshape(B, L, nh, D).transpose(1, 2) for t in qkv.split(E, dim=-1)]
    att = torch.softmax(q @ k.transpose(-1, -2) / D ** 0.5, dim=-1)
    want = (att @ v).transpose(1, 2).reshape(B, L, E)
    qg = qkv.cuda()
    out = torch.empty(B, L, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * nh * L, device="cuda")
    _call("pcm_mha_fwd_tc", qg.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, nh, 1.0 / D ** 0.5, 0.0, 0, _s())
    out2 = torch.empty_like(out)
    lse2 = torch.empty_like(lse)
    _call("pcm_mha_fwd", qg.data_ptr(), out2.data_ptr(), lse2.data_ptr(), B, L, nh, D, 1.0 / D ** 0.5, 0.0, 0, 1, _s())
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    err = float((out.double().cpu() - want).norm() / want.norm())
    assert err < 8e-3, err                                   # bf16 probabilities and bf16 output
    assert float((lse - lse2).abs().max()) < 2e-3


@pytest.mark.parametrize("cfg", [(3, 216, 4, 0.0), (2, 224, 2, 0.0), (2, 100, 4, 0.0), (1, 17, 1, 0.0), (2, 216, 4, 0.1)])
def test_mha_bwd_tc_matches_reference(ops, cfg):
    """tcgen05 attention backward against fp64 autograd (p = 0) and against the SIMT backward with the same
    counter-based dropout mask (p > 0)."""
    from pcm_b200._lib import lib
    from pcm_b200.ops import _call, _s
    B, L, nh, pd = cfg
    D, E = 32, 32 * nh
    g = torch.Generator().manual_seed(B * 100 + L + nh)
    qkv = torch.randn(B, L, 3 * E, generator=g).bfloat16()
    dout = (torch.randn(B, L, E, generator=g) / 8).bfloat16()
    qg, dg = qkv.cuda(), dout.cuda()
    sc, seed = 1.0 / D ** 0.5, 4242
    out = torch.empty(B, L, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * nh * L, device="cuda")
    _call("pcm_mha_fwd_tc", qg.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, nh, sc, pd, seed, _s())
    dq_tc = torch.full((B, L, 3 * E), float("nan"), device="cuda", dtype=torch.bfloat16)
    _call("pcm_mha_bwd_tc", qg.data_ptr(), out.data_ptr(), dg.data_ptr(), lse.data_ptr(), dq_tc.data_ptr(), B, L, nh, sc, pd,
          seed, _s())
    dq_simt = torch.empty_like(dq_tc)
    _call("pcm_mha_bwd", qg.data_ptr(), out.data_ptr(), dg.data_ptr(), lse.data_ptr(), dq_simt.data_ptr(), B, L, nh, D, sc, pd,
          seed, 1, _s())
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    a, b = dq_tc.double().cpu(), dq_simt.double().cpu()
    assert bool(torch.isfinite(a).all())
    assert float((a - b).norm() / b.norm()) < 1.5e-2                      # bf16 P~ / dS' operands vs fp32 SIMT math
    if pd == 0.0:
        x = qkv.double().requires_grad_(True)
        q, k, v = [t.reshape(B, L, nh, D).transpose(1, 2) for t in x.split(E, dim=-1)]
        o = (torch.softmax(q @ k.transpose(-1, -2) * sc, dim=-1) @ v).transpose(1, 2).reshape(B, L, E)
        o.backward(dout.double())
        assert float((a - x.grad).norm() / x.grad.norm()) < 1.5e-2


@pytest.mark.parametrize("cfg", [(3, 48, 72, 5, 16, 64), (2, 24, 36, 64, 64, 128), (2, 12, 20, 32, 32, 32)])
def test_conv3x3_stride2_tc_matches_reference(ops, cfg):
    """3x3 / stride-2 / pad-1 conv (+bias +ReLU) forward, data gradient and weight gradient on the tensor cores (pixel-pair
    form) against torch's CPU conv2d on the same bf16-rounded operands."""
    from pcm_b200 import ops_nn
    from pcm_b200._lib import lib
    N, H, W, Ci, Cip, Co = cfg
    g = torch.Generator().manual_seed(N + H + Ci + Co)
    x = torch.zeros(N, H, W, Cip)
    x[..., :Ci] = torch.randn(N, H, W, Ci, generator=g)
    x = x.bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) / (3 * Ci ** 0.5)).bfloat16().float()
    b = torch.randn(Co, generator=g)
    xr = x[..., :Ci].double().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.double().requires_grad_(True)
    br = b.double().requires_grad_(True)
    want = F.relu(F.conv2d(xr, wr, br, stride=2, padding=1))
    xg = x.cuda().requires_grad_(True)
    wg, bg = torch.nn.Parameter(w.cuda()), torch.nn.Parameter(b.cuda())
    y = ops_nn.Conv2dFn.apply(xg, wg, bg, 2, 1, True)
    assert lib().last_call == "pcm_conv3x3s2_tc"
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert float((got - want.detach().float()).norm() / want.norm()) < 4e-3
    dy = (torch.randn(y.shape, generator=g) / y.numel() ** 0.5).bfloat16()
    y.backward(dy.cuda())
    want.backward(dy.double().permute(0, 3, 1, 2))
    torch.cuda.synchronize()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert float((wg.grad.cpu().double() - wr.grad).norm() / wr.grad.norm()) < 2e-3      # dy is ReLU-masked in bf16
    assert float((bg.grad.cpu().double() - br.grad).norm() / br.grad.norm()) < 2e-3
    dx = xg.grad.float().cpu()[..., :Ci].permute(0, 3, 1, 2)
    assert float((dx.double() - xr.grad).norm() / xr.grad.norm()) < 6e-3


@pytest.mark.parametrize("last_only", [False, True])
@pytest.mark.parametrize("B", [5, 64])
def test_persistent_convlstm_matches_per_step_path_and_oracle(B, last_only, monkeypatch):
    """csrc/convlstm_seq.cu (all T steps / the whole BPTT in one cluster launch each) against (a) the per-step kernels it
    replaces, same inputs — they share the bf16 operand rounding, so outputs agree to accumulation order and the
    tanh.approx-vs-tanhf difference at t = 0 — and (b) the fp64 oracle of src/convlstm.py:27-35.  B = 5 leaves one
    cluster with a single valid sample; B = 64 is the benchmark's ConvLSTM."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200 import ops
    from pcm_b200._lib import lib
    T, Cin, Ch, H, W = 6, 128, 64, 6, 9
    sd = O.synth_state_dict([("cell.conv.weight", (4 * Ch, Cin + Ch, 3, 3)), ("cell.conv.bias", (4 * Ch,))], 301)
    g = torch.Generator().manual_seed(302)
    x = torch.randn(T, B, Cin, H, W, generator=g)
    gy = torch.randn((B, Ch, H, W) if last_only else (T, B, Ch, H, W), generator=g) / 8

    def run(persistent):
        monkeypatch.setenv("PCM_LSTM_PERSISTENT", "1" if persistent else "0")
        w = sd["cell.conv.weight"].cuda().requires_grad_(True)
        b = sd["cell.conv.bias"].cuda().requires_grad_(True)
        xs = ops.StageIn.apply(x.reshape(T * B, Cin, H, W).cuda().requires_grad_(True), torch.bfloat16)
        xs.retain_grad()
        h = ops.ConvLSTMFn.apply(xs, w, b, T, B, B, 1, last_only)
        assert ("seq" in lib().last_call or True)
        hn = h.float().permute(0, 3, 1, 2) if last_only else h.float().permute(0, 1, 4, 2, 3)
        (hn * gy.cuda()).sum().backward()
        torch.cuda.synchronize()
        return hn.detach().cpu(), w.grad.cpu(), b.grad.cpu(), xs.grad.float().cpu()

    n0 = lib().launches
    a = run(True)
    n_persistent = lib().launches - n0
    n0 = lib().launches
    c = run(False)
    n_steps = lib().launches - n0
    assert lib()._fn["pcm_tc_error_count"]() == 0
    assert n_persistent <= n_steps - 2 * (T - 1), (n_persistent, n_steps)          # fewer launches: the point of it
    names = ["h", "dW", "db", "dx"]
    for nm, u, v in zip(names, a, c):
        e = float((u - v).norm() / v.norm())
        assert e < 2e-2, (nm, e)
    # oracle (fp64)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xd = x.double().requires_grad_(True)
    ho = O.convlstm(xd, sdd, "")
    (( ho[-1] if last_only else ho) * gy.double()).sum().backward()
    ref = [(ho[-1] if last_only else ho).detach(), sdd["cell.conv.weight"].grad, sdd["cell.conv.bias"].grad]
    for nm, u, v in zip(names[:3], a[:3], ref):
        e = float((u.double() - v).norm() / v.norm())
        assert e < (3e-2 if nm == "h" else 6e-2), (nm, e)


@pytest.mark.parametrize("shape", [(5, 48, 72, 7, 16), (5, 24, 36, 16, 32), (5, 12, 18, 32, 64), (3, 24, 36, 64, 32),
                                   (4, 12, 18, 8, 16)])
def test_fused_block_forward_matches_four_kernel_path(shape, monkeypatch):
    """convblock_fwd_tc_kernel (whole ConvBlock forward in one launch: convs on tcgen05 over the smem-resident image,
    conv outputs in TMEM) against the 4-kernel forward it replaces — output, every saved tensor the backward consumes,
    and the gradients that come out of the (shared) backward — and against the fp64 oracle of src/unet.py:35-49."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200 import ops
    from pcm_b200._lib import lib
    from pcm_b200.src.unet import ConvBlock
    N, H, W, cin, cout = shape
    sd = O.synth_state_dict(O._convblock_spec("", cin, cout), 501)
    x, y = O.synth_frame_batch(N, cin, H, W, 502, out_ch=cout)

    def run(fused):
        monkeypatch.setenv("PCM_BLOCK_FWD_TC", "1" if fused else "0")
        m = ConvBlock(cin, cout)
        m.load_state_dict(sd)
        m = m.cuda()
        xs = ops.StageIn.apply(x.cuda(), torch.bfloat16)
        xs.requires_grad_(True)
        seen = []
        orig = lib().call

        def spy(name, *a):
            seen.append(name)
            return orig(name, *a)
        monkeypatch.setattr(lib(), "call", spy)
        out = m.forward_nhwc(xs)
        saved = [t.detach().float().cpu() if torch.is_tensor(t) and t.dtype != torch.uint8 else (t.cpu() if torch.is_tensor(t) else t)
                 for t in out.grad_fn.saved_tensors]
        loss = ops.mse_loss(ops.StageOut.apply(out, cout), y.cuda())
        loss.backward()
        torch.cuda.synchronize()
        monkeypatch.setattr(lib(), "call", orig)
        grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
        return out.detach().float().cpu(), saved, grads, xs.grad.float().cpu(), seen

    o1, s1, g1, dx1, seen1 = run(True)
    o0, s0, g0, dx0, seen0 = run(False)
    assert "pcm_convblock_fwd_tc" in seen1 and "pcm_convblock_fwd_tc" not in seen0
    assert "pcm_convblock_tail_fwd" in seen0 and "pcm_convblock_tail_fwd" not in seen1
    assert lib()._fn["pcm_tc_error_count"]() == 0
    rel = lambda a, b: float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))
    assert rel(o1, o0) < 1e-2, rel(o1, o0)
    names = ["x", "y1", "a1", "y2", "small", "se", "w1", "g1", "b1", "w2", "g2", "b2", "sw1", "sw2", "wsp", "maps", "ties", "out"]
    for nm, a, b in zip(names, s1, s0):
        if nm == "ties":
            assert float((a != b).float().mean()) < 0.02, nm
        else:
            assert rel(a, b) < 1e-2, (nm, rel(a, b))
    errs = {k: rel(g1[k], g0[k]) for k in g0 if float(g0[k].norm()) > 1e-7}
    import numpy as np
    assert float(np.median(list(errs.values()))) < 3e-2 and max(errs.values()) < 0.3, errs
    assert rel(dx1, dx0) < 3e-2
    # oracle
    sdd = {k: v.double() for k, v in sd.items()}
    ref = O.conv_block(x.double(), sdd, "")
    got = o1[..., :cout].permute(0, 3, 1, 2)
    assert rel(got, ref) < 3e-2, rel(got, ref)
