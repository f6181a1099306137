"""Host-side operators for the SimpleCNN and CNNTransformer variants (reference src/models.py:44-123,
src/cnn_transformer.py:5-54): general Conv2d / ConvTranspose2d(k2,s2) / Linear with bias, BatchNorm2d (training
statistics, running-stat update), ReLU, Dropout / Dropout2d, residual + LayerNorm and multi-head attention.

Same conventions as ops.py: NHWC activations in the compute dtype, fp32 nn.Parameters in reference shapes, weight
gradients accumulated by the kernels into `param.main_grad` when the trainer provides it.  Tensor-core shapes go to
the tcgen05 kernels (pcm_conv3x3_tc / pcm_conv1x1_tc / pcm_wgrad*_tc), everything else (fp32 parity path, thin or
strided layers) to the general SIMT gather kernels.  No CPU / eager fallback."""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from .ops import _DT, _call, _grad_buf, _p, _require_cuda, _s, channel_sum, conv_gather, conv_wgrad, pack_weight

BN_EPS = 1e-5
LN_EPS = 1e-5

_seed_counter = [0]


def next_seed() -> int:
    """Per-call dropout seed: torch's initial seed (so torch.manual_seed makes runs reproducible) + a counter."""
    _seed_counter[0] += 1
    return (int(torch.initial_seed()) * 1000003 + _seed_counter[0]) & 0x7FFFFFFFFFFFFFFF


# ------------------------------------------------------------------------------------------------
# stride-1 "same" convolution with K in {1, 3}: tensor cores with Cout / Cin splitting, SIMT otherwise
# ------------------------------------------------------------------------------------------------
def _tc_in_ok(dtype, Sc):
    return dtype == torch.bfloat16 and (Sc in (16, 32) or (Sc >= 64 and Sc % 64 == 0))


def conv_same(src, wk, N, H, W, Sc, Dc, K, bias=None, relu=False):
    """dst (N,H,W,Dc) = conv_KxK(src (N,H,W,Sc), wk [K*K][Dc][Sc]) (+bias)(relu), stride 1, pad K//2."""
    dt = src.dtype
    if not (_tc_in_ok(dt, Sc) and K in (1, 3) and Dc % 16 == 0):
        return conv_gather(src, wk, N, H, W, Sc, H, W, Dc, K, K, 1, K // 2, 0, bias=bias, relu=relu)
    dst = torch.empty((N, H, W, Dc), device=src.device, dtype=dt)
    # at most 256 accumulator columns per launch (double-buffered TMEM): split the output channels
    nsplit = (Dc + 255) // 256
    step = ((Dc // nsplit) + 15) // 16 * 16
    taps = K * K
    c0 = 0
    while c0 < Dc:
        cs = min(step, Dc - c0)
        if nsplit == 1:
            wk_s = wk
        else:
            wk_s = wk[:, c0:c0 + cs, :].contiguous()            # D2D copy of a weight slice (plumbing)
        b_ptr = 0 if bias is None else bias.data_ptr() + 4 * c0
        if K == 3 and not relu:
            _call("pcm_conv3x3_tc", src.data_ptr(), H * W * Sc, Sc, H, W, Sc, dst.data_ptr() + c0 * dst.element_size(),
                  H * W * Dc, Dc, cs, wk_s.data_ptr(), b_ptr, N, 0, 0, _s())
        elif K == 1:
            _call("pcm_conv1x1_tc", src.data_ptr(), H * W * Sc, Sc, H, W, Sc, dst.data_ptr() + c0 * dst.element_size(),
                  H * W * Dc, Dc, cs, wk_s.data_ptr(), b_ptr, N, 0, 0, int(relu), _s())
        else:
            raise RuntimeError("conv_same: fused ReLU is only wired for K=1")
        c0 += cs
    return dst


def wgrad_same(dy, x, dw, N, H, W, Co, Ci, Ci_real, K):
    """dw (Co, Ci_real, K, K) fp32 += sum_p dy(p,co) x(p+tap,ci) for the stride-1 same conv."""
    KK = K * K
    tc = (dy.dtype == torch.bfloat16 and K in (1, 3) and (Co in (16, 32, 64) or Co % 128 == 0)
          and (Ci in (16, 32, 64, 128, 192, 256) or (Ci > 256 and Ci % 256 == 0)) and W + 2 <= 256 and H + 2 <= 256)
    if not tc:
        conv_wgrad(dy, x, dw, Ci_real * KK, KK, 1, N, H, W, Co, Co, H, W, Ci, Ci_real, K, K, 1, K // 2)
        return
    c0 = 0
    while c0 < Ci:
        cs = min(256, Ci - c0)
        cr = max(0, min(cs, Ci_real - c0))
        if cr > 0:
            if K == 3:
                _call("pcm_wgrad3x3_tc", dy.data_ptr(), H * W * Co, Co, Co, Co, x.data_ptr() + c0 * x.element_size(),
                      H * W * Ci, Ci, cs, cr, dw.data_ptr() + 4 * c0 * KK, Ci_real * KK, KK, 1, N, H, W, _s())
            else:
                _call("pcm_wgrad1x1_tc", dy.data_ptr(), H * W * Co, Co, Co, Co, x.data_ptr() + c0 * x.element_size(),
                      H * W * Ci, Ci, cs, cr, dw.data_ptr() + 4 * c0, Ci_real, 1, N, H, W, _s())
        c0 += cs


# ------------------------------------------------------------------------------------------------
# 3x3 / stride-2 / pad-1 convolution in "pixel pair" form on the tensor cores (csrc/conv_tc.cu mode 2)
# ------------------------------------------------------------------------------------------------
_S2_IDX = {}


def _s2_indices(Co, Ci, Cip, dev):
    """Gather tables (built once per shape) from the flat (Co, Ci, 3, 3) weight [+ one trailing zero] into
    fwd  [6][Co][2*Cip]   : tap kh*2 + dwp, channel pw*Cip + c  <- w[co][c][kh][2*dwp + pw - 1]
    dgrad[4][4*Cip][Co]   : tap oh*2 + ow, row (ph*2+pw)*Cip + c <- w[co][c][kh(ph,oh)][kw(pw,ow)]
    and from the flat packed weight gradient [6][Co][2*Cip] back to (Co, Ci, 3, 3)."""
    key = (Co, Ci, Cip, str(dev))
    if key in _S2_IDX:
        return _S2_IDX[key]
    zero = Co * Ci * 9
    widx = lambda co, c, kh, kw: ((co * Ci + c) * 3 + kh) * 3 + kw
    fwd = torch.full((6, Co, 2 * Cip), zero, dtype=torch.long)
    fold = torch.zeros((Co, Ci, 3, 3), dtype=torch.long)
    co = torch.arange(Co).reshape(Co, 1)
    c = torch.arange(Ci).reshape(1, Ci)
    for kh in range(3):
        for dwp in range(2):
            for pw in range(2):
                kw = 2 * dwp + pw - 1
                if kw < 0:
                    continue
                fwd[kh * 2 + dwp, :, pw * Cip: pw * Cip + Ci] = widx(co, c, kh, kw)
                fold[:, :, kh, kw] = ((kh * 2 + dwp) * Co + co) * (2 * Cip) + pw * Cip + c
    dgr = torch.full((4, 4 * Cip, Co), zero, dtype=torch.long)
    tap_of = {(0, 0): 1, (1, 1): 0, (1, 0): 2}                  # (parity, offset) -> kernel index; (0, 1) unused
    cc = torch.arange(Ci).reshape(Ci, 1)
    oo = torch.arange(Co).reshape(1, Co)
    for ph in range(2):
        for oh in range(2):
            if (ph, oh) not in tap_of:
                continue
            for pw in range(2):
                for ow in range(2):
                    if (pw, ow) not in tap_of:
                        continue
                    q = ph * 2 + pw
                    dgr[oh * 2 + ow, q * Cip: q * Cip + Ci, :] = widx(oo, cc, tap_of[(ph, oh)], tap_of[(pw, ow)])
    out = (fwd.reshape(-1).to(dev), dgr.reshape(-1).to(dev), fold.reshape(-1).to(dev))
    _S2_IDX[key] = out
    return out


def _s2_supported(x, w, stride, pad):
    N, H, W, Cip = x.shape
    Co, Ci, K, _ = w.shape
    return (x.dtype == torch.bfloat16 and K == 3 and stride == 2 and pad == 1 and H % 2 == 0 and W % 2 == 0 and
            Cip in (16, 32, 64) and Co % 16 == 0 and Co <= 256 and (Co in (16, 32, 64) or Co % 128 == 0) and
            (Co * Ci * 9) % 8 == 0 and W // 2 <= 256 and H // 2 <= 256 and os.environ.get("PCM_S2_TC", "1") != "0")


def _gather_weight(w, idx, shape, dtype):
    """Layout gather of a weight tensor (index plumbing: every output element is one input element or zero)."""
    flat = torch.cat([w.detach().reshape(-1), w.new_zeros(1)])
    return flat[idx].reshape(shape).to(dtype)


class Conv2dFn(torch.autograd.Function):
    """nn.Conv2d(Ci, Co, K, stride, padding=K//2 (stride 1) or 1 (K=3, stride 2), bias) on NHWC, optional fused ReLU
    (src/models.py:47,50,57,90,108; src/cnn_transformer.py:10,12)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, relu):
        _require_cuda(x, "activation")
        x = x.contiguous()
        N, H, W, Cip = x.shape
        Co, Ci, K, _ = w.shape
        assert Ci <= Cip and Co % 8 == 0
        dt = x.dtype
        Ho, Wo = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
        if _s2_supported(x, w, stride, pad):
            fwd_idx, _, _ = _s2_indices(Co, Ci, Cip, x.device)
            wk = _gather_weight(w, fwd_idx, (6, Co, 2 * Cip), dt)
            y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=dt)
            _call("pcm_conv3x3s2_tc", x.data_ptr(), H * W * Cip, Cip, Ho, Wo, y.data_ptr(), Ho * Wo * Co, Co, Co, wk.data_ptr(),
                  _p(b), N, int(relu), _s())
            ctx.save_for_backward(x, w, b, y if relu else None)
            ctx.cfg = (stride, pad, relu, Ho, Wo)
            return y
        wk = pack_weight(w, Ci * K * K, K * K, 1, Co, Ci, K * K, dt, Ip=Cip)
        if stride == 1 and pad == K // 2:
            y = conv_same(x, wk, N, H, W, Cip, Co, K, bias=b, relu=relu)
        else:
            y = conv_gather(x, wk, N, H, W, Cip, Ho, Wo, Co, K, K, stride, pad, 0, bias=b, relu=relu)
        ctx.save_for_backward(x, w, b, y if relu else None)
        ctx.cfg = (stride, pad, relu, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, y = ctx.saved_tensors
        stride, pad, relu, Ho, Wo = ctx.cfg
        N, H, W, Cip = x.shape
        Co, Ci, K, _ = w.shape
        dt = x.dtype
        dy = dy.contiguous()
        if relu:
            dz = torch.empty_like(dy)
            _call("pcm_relu_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), _DT[dt], _s())
            dy = dz
        gw, rw = _grad_buf(w)
        KK = K * K
        s2 = _s2_supported(x, w, stride, pad)
        if s2:
            _, dgr_idx, fold_idx = _s2_indices(Co, Ci, Cip, x.device)
        rb = None
        with ops.side_stream(dy, x):            # weight / bias gradients beside the data-gradient chain
            if s2:
                tmp = torch.zeros((6, Co, 2 * Cip), device=x.device, dtype=torch.float32)
                _call("pcm_wgrad3x3s2_tc", dy.data_ptr(), Ho * Wo * Co, Co, Co, x.data_ptr(), H * W * Cip, Cip, tmp.data_ptr(),
                      2 * Cip, 1, Co * 2 * Cip, N, Ho, Wo, _s())
                folded = tmp.reshape(-1)[fold_idx]                              # layout gather back to (Co, Ci, 3, 3)
                _call("pcm_add", gw.data_ptr(), folded.data_ptr(), gw.data_ptr(), gw.numel(), 0, _s())
            elif stride == 1 and pad == K // 2:
                wgrad_same(dy, x, gw, N, H, W, Co, Cip, Ci, K)
            else:
                conv_wgrad(dy, x, gw, Ci * KK, KK, 1, N, Ho, Wo, Co, Co, H, W, Cip, Ci, K, K, stride, pad)
            if b is not None:
                gb, rb = _grad_buf(b)
                channel_sum(dy, gb, N, Ho * Wo, Co, Co)
        dx = None
        if ctx.needs_input_grad[0] and s2:
            wkd = _gather_weight(w, dgr_idx, (4, 4 * Cip, Co), dt)
            dx = torch.empty_like(x)
            _call("pcm_conv3x3s2_dgrad_tc", dy.data_ptr(), Ho * Wo * Co, Co, Ho, Wo, Co, dx.data_ptr(), H * W * Cip, Cip, Cip,
                  wkd.data_ptr(), N, _s())
        elif ctx.needs_input_grad[0]:
            if stride == 1 and pad == K // 2:
                # data gradient == forward conv of dy with flipped taps and transposed channels
                wkt = pack_weight(w, KK, Ci * KK, -1, Ci, Co, KK, dt, offset=KK - 1, Op=Cip)
                dx = conv_same(dy, wkt, N, H, W, Co, Cip, K)
            else:
                wkt = pack_weight(w, KK, Ci * KK, 1, Ci, Co, KK, dt, Op=Cip)      # wk[tap][ci][co]
                dx = conv_gather(dy, wkt, N, Ho, Wo, Co, H, W, Cip, K, K, stride, pad, 1)
        return dx, rw, rb, None, None, None


class ConvT2x2Fn(torch.autograd.Function):
    """nn.ConvTranspose2d(Ci, Co, kernel_size=2, stride=2) + bias (+ReLU) on NHWC (src/cnn_transformer.py:36,38)."""

    @staticmethod
    def forward(ctx, x, wt, bt, relu):
        x = x.contiguous()
        B, h, w, Ci = x.shape
        Co = wt.shape[1]
        assert wt.shape[0] == Ci and Co % 8 == 0
        dt = x.dtype
        y = torch.empty((B, 2 * h, 2 * w, Co), device=x.device, dtype=dt)
        ops.convT2x2_fwd(x, wt, bt, B, h, w, Ci, Co, y, 4 * h * w * Co, Co, relu=relu)
        ctx.save_for_backward(x, wt, bt, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wt, bt, y = ctx.saved_tensors
        B, h, w, Ci = x.shape
        Co = wt.shape[1]
        dt = x.dtype
        dy = dy.contiguous()
        if ctx.relu:
            dz = torch.empty_like(dy)
            _call("pcm_relu_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), _DT[dt], _s())
            dy = dz
        gwt, rwt = _grad_buf(wt)
        gbt, rbt = _grad_buf(bt)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.convT2x2_dgrad(dy, wt, B, h, w, Ci, Co, 4 * h * w * Co, Co)
        with ops.side_stream(dy, x):
            ops.convT2x2_wgrad(x, dy, gwt, B, h, w, Ci, Co, 4 * h * w * Co, Co)
            channel_sum(dy, gbt, B, 4 * h * w, Co, Co)
        return dx, rwt, rbt, None


# ------------------------------------------------------------------------------------------------
# BatchNorm2d (+ residual add + ReLU)
# ------------------------------------------------------------------------------------------------
class BatchNormFn(torch.autograd.Function):
    """y = [relu]( BN(x) [+ res] ).  training: batch statistics over (N,H,W), running statistics updated in place
    like nn.BatchNorm2d (momentum 0.1, unbiased running variance); eval: running statistics.
    Reference: src/models.py:48,51,57,62-71,91,109."""

    @staticmethod
    def forward(ctx, x, gamma, beta, res, relu, training, running_mean, running_var, nbt, momentum):
        _require_cuda(x, "activation")
        x = x.contiguous()
        C = x.shape[-1]
        R = x.numel() // C
        d, st = _DT[x.dtype], _s()
        if training:
            sums = torch.zeros(C * 2, device=x.device, dtype=torch.float64)     # (sum, sum of squares) in double
            _call("pcm_bn_stats", x.data_ptr(), sums.data_ptr(), R, C, d, st)
            if running_mean is not None:
                _call("pcm_bn_update_running", sums.data_ptr(), running_mean.data_ptr(), running_var.data_ptr(),
                      _p(nbt), R, C, momentum, st)
        else:
            # the kernels derive mean/var from (sum, sum of squares) / R: encode the running statistics that way
            # (a C-element host-side staging of buffers, not activation arithmetic)
            rm, rv = running_mean.double(), running_var.double()
            sums = (torch.stack([rm, rv + rm * rm], dim=1) * float(R)).reshape(-1).contiguous()
        y = torch.empty_like(x)
        if res is not None:
            res = res.contiguous()
        _call("pcm_bn_apply_fwd", x.data_ptr(), sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(res), y.data_ptr(),
              R, C, BN_EPS, int(relu), d, st)
        ctx.save_for_backward(x, gamma, beta, sums, y if relu else None)
        ctx.cfg = (relu, res is not None, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, sums, y = ctx.saved_tensors
        relu, has_res, training = ctx.cfg
        if not training:
            raise RuntimeError("pcm_b200 BatchNorm: backward through eval-mode statistics is not supported")
        C = x.shape[-1]
        R = x.numel() // C
        d, st = _DT[x.dtype], _s()
        dy = dy.contiguous()
        gg, rg = _grad_buf(gamma)
        gb, rb = _grad_buf(beta)
        dsum = torch.zeros(C * 2, device=x.device, dtype=torch.float32)
        _call("pcm_bn_bwd_reduce", dy.data_ptr(), _p(y), x.data_ptr(), sums.data_ptr(), dsum.data_ptr(), R, C, BN_EPS, d, st)
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (has_res and ctx.needs_input_grad[3]) else None
        _call("pcm_bn_bwd_apply", dy.data_ptr(), _p(y), x.data_ptr(), sums.data_ptr(), gamma.data_ptr(), dsum.data_ptr(),
              dx.data_ptr(), _p(dres), gg.data_ptr(), gb.data_ptr(), R, C, BN_EPS, d, st)
        return dx, rg, rb, dres, None, None, None, None, None, None


def batch_norm(x, bn: torch.nn.BatchNorm2d, res=None, relu=False):
    """Apply an nn.BatchNorm2d container's parameters/buffers with the pcm kernels."""
    training = bn.training or bn.running_mean is None
    return BatchNormFn.apply(x, bn.weight, bn.bias, res, relu, training, bn.running_mean, bn.running_var,
                             bn.num_batches_tracked, float(bn.momentum if bn.momentum is not None else 0.1))


class AddFn(torch.autograd.Function):
    """a + b (same shape) — gradient fan-in where a residual branch rejoins (src/models.py:70)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        _call("pcm_add", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _DT[a.dtype], _s())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


_PENDING_ADD = {}      # data_ptr of a gradient handed on by a lazy fork -> its second addend (consumed by AddLayerNormFn.backward)


class ForkFn(torch.autograd.Function):
    """x -> (x, x) for a tensor consumed by two branches (residual connections): the two incoming gradients are
    summed by pcm_add instead of autograd's own accumulation kernel.  lazy=True (ONLY when x is the output of
    AddLayerNormFn, whose backward reads the pair): the sum is not materialised — g1 is handed on and g2 parked in
    _PENDING_ADD under g1's address; pcm_layernorm_bwd adds the two while loading (one launch and one tensor pass fewer
    per residual connection)."""

    @staticmethod
    def forward(ctx, x, lazy=False):
        ctx.lazy = bool(lazy)
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None:
            return g2, None
        if g2 is None:
            return g1, None
        g1, g2 = g1.contiguous(), g2.contiguous()
        if ctx.lazy:
            _PENDING_ADD[g1.data_ptr()] = g2
            return g1, None
        out = torch.empty_like(g1)
        _call("pcm_add", g1.data_ptr(), g2.data_ptr(), out.data_ptr(), g1.numel(), _DT[g1.dtype], _s())
        return out, None


def fork(x, lazy=False):
    """lazy=True only for the output of AddLayerNormFn (see ForkFn)."""
    return ForkFn.apply(x, lazy) if x.requires_grad else (x, x)


class Dropout2dFn(torch.autograd.Function):
    """nn.Dropout2d(p) (src/models.py:103,120): whole channels of each sample are zeroed, survivors scaled 1/(1-p)."""

    @staticmethod
    def forward(ctx, x, p, seed):
        x = x.contiguous()
        N, H, W, C = x.shape
        mask = torch.empty(N * C, device=x.device, dtype=torch.float32)
        _call("pcm_dropout_mask", mask.data_ptr(), N * C, p, seed, _s())
        y = torch.empty_like(x)
        _call("pcm_scale_channels", x.data_ptr(), mask.data_ptr(), 0, y.data_ptr(), N, H * W, C, _DT[x.dtype], _s())
        ctx.save_for_backward(mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        (mask,) = ctx.saved_tensors
        dy = dy.contiguous()
        N, H, W, C = dy.shape
        dx = torch.empty_like(dy)
        _call("pcm_scale_channels", dy.data_ptr(), mask.data_ptr(), 0, dx.data_ptr(), N, H * W, C, _DT[dy.dtype], _s())
        return dx, None, None


class DropoutFn(torch.autograd.Function):
    """nn.Dropout(p) with a counter-based mask: backward re-applies the same call to the gradient."""

    @staticmethod
    def forward(ctx, x, p, seed):
        x = x.contiguous()
        y = torch.empty_like(x)
        _call("pcm_dropout", x.data_ptr(), y.data_ptr(), x.numel(), p, seed, _DT[x.dtype], _s())
        ctx.cfg = (p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        _call("pcm_dropout", dy.data_ptr(), dx.data_ptr(), dy.numel(), p, seed, _DT[dy.dtype], _s())
        return dx, None, None


def dropout(x, p: float, training: bool):
    if not training or p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p), next_seed())


# ------------------------------------------------------------------------------------------------
# transformer encoder layer pieces (src/cnn_transformer.py:25-31)
# ------------------------------------------------------------------------------------------------
class AddPosFn(torch.autograd.Function):
    """x (B, L, E) + pos_embedding (1, L, E) (src/cnn_transformer.py:48)."""

    @staticmethod
    def forward(ctx, x, pos):
        x = x.contiguous()
        B = x.shape[0]
        LE = x.numel() // B
        y = torch.empty_like(x)
        _call("pcm_add_bcast", x.data_ptr(), pos.data_ptr(), y.data_ptr(), x.numel(), LE, _DT[x.dtype], _s())
        ctx.save_for_backward(pos)
        ctx.dims = (B, LE)
        return y

    @staticmethod
    def backward(ctx, dy):
        (pos,) = ctx.saved_tensors
        B, LE = ctx.dims
        dy = dy.contiguous()
        gp, rp = _grad_buf(pos)
        with ops.side_stream(dy):
            _call("pcm_batch_sum", dy.data_ptr(), gp.data_ptr(), B, LE, _DT[dy.dtype], _s())   # dpos = sum_b dy[b]
        return dy, rp


class LinearFn(torch.autograd.Function):
    """nn.Linear(K, N) (+ReLU) on a token matrix x (B, L, K): the 1x1 tensor-core path with the tokens of one
    sample as a 1 x L image."""

    @staticmethod
    def forward(ctx, x, w, b, relu, drop_p=0.0, seed=0):
        """drop_p > 0: nn.Dropout(p) on the output as well (linear1 -> ReLU -> dropout of the transformer FFN) — in the
        GEMM's store epilogue on the tensor-core path (pcm_conv1x1_drop_tc), by pcm_dropout otherwise; the same mask either
        way.  The backward then needs only the saved output: dz = y > 0 ? dy / (1 - p) : 0."""
        x = x.contiguous()
        B, L, K = x.shape
        N = w.shape[0]
        dt = x.dtype
        wk = pack_weight(w, K, 1, 0, N, K, 1, dt)                               # [1][N][K]
        drop_p = float(drop_p)
        if drop_p > 0.0 and relu and _tc_in_ok(dt, K) and N % 16 == 0 and N <= 256:
            y = torch.empty((B, L, N), device=x.device, dtype=dt)
            _call("pcm_conv1x1_drop_tc", x.data_ptr(), L * K, K, 1, L, K, y.data_ptr(), L * N, N, N, wk.data_ptr(), _p(b), B, 1,
                  drop_p, int(seed), _s())
        else:
            y = conv_same(x, wk, B, 1, L, K, N, 1, bias=b, relu=relu).reshape(B, L, N)
            if drop_p > 0.0:
                yd = torch.empty_like(y)
                _call("pcm_dropout", y.data_ptr(), yd.data_ptr(), y.numel(), drop_p, int(seed), _DT[dt], _s())
                y = yd
        ctx.save_for_backward(x, w, b, y if (relu or drop_p > 0.0) else None)
        ctx.relu = relu
        ctx.drop = (drop_p, int(seed))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, y = ctx.saved_tensors
        B, L, K = x.shape
        N = w.shape[0]
        dt = x.dtype
        dy = dy.contiguous()
        drop_p, seed = ctx.drop
        if ctx.relu and drop_p > 0.0:
            dz = torch.empty_like(dy)
            _call("pcm_relu_bwd_scaled", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), 1.0 / (1.0 - drop_p),
                  _DT[dt], _s())
            dy = dz
        elif ctx.relu:
            dz = torch.empty_like(dy)
            _call("pcm_relu_bwd", dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), _DT[dt], _s())
            dy = dz
        elif drop_p > 0.0:
            dz = torch.empty_like(dy)
            _call("pcm_dropout", dy.data_ptr(), dz.data_ptr(), dy.numel(), drop_p, seed, _DT[dt], _s())
            dy = dz
        gw, rw = _grad_buf(w)
        gb, rb = _grad_buf(b)
        with ops.side_stream(dy, x):            # parameter gradients: nothing in the backward chain waits for them
            wgrad_same(dy, x, gw, B, 1, L, N, K, K, 1)
            channel_sum(dy, gb, B, L, N, N)
        dx = None
        if ctx.needs_input_grad[0]:
            wkt = pack_weight(w, 1, K, 0, K, N, 1, dt)                          # [1][K][N] = W^T
            dx = conv_same(dy, wkt, B, 1, L, N, K, 1).reshape(B, L, K)
        return dx, rw, rb, None, None, None


class AddLayerNormFn(torch.autograd.Function):
    """LayerNorm(a + dropout_p(b)) over the last dim (post-norm residual, eps 1e-5).  The sub-layer dropout
    (`dropout1` / `dropout2` of nn.TransformerEncoderLayer) is applied to b inside the kernel, and the backward kernel
    emits both gradients (of a, and of b = the same mask applied to it)."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, drop_p=0.0, seed=0):
        a, b = a.contiguous(), b.contiguous()
        E = a.shape[-1]
        M = a.numel() // E
        s = torch.empty_like(a)
        y = torch.empty_like(a)
        stat = torch.empty(M * 2, device=a.device, dtype=torch.float32)
        _call("pcm_add_layernorm_fwd", a.data_ptr(), b.data_ptr(), gamma.data_ptr(), beta.data_ptr(), s.data_ptr(),
              y.data_ptr(), stat.data_ptr(), M, E, LN_EPS, float(drop_p), int(seed), _DT[a.dtype], _s())
        ctx.save_for_backward(s, stat, gamma, beta)
        ctx.drop = (float(drop_p), int(seed))
        return y

    @staticmethod
    def backward(ctx, dy):
        s, stat, gamma, beta = ctx.saved_tensors
        drop_p, seed = ctx.drop
        E = s.shape[-1]
        M = s.numel() // E
        dy = dy.contiguous()
        gg, rg = _grad_buf(gamma)
        gb, rb = _grad_buf(beta)
        ds = torch.empty_like(s)
        db = torch.empty_like(s) if drop_p > 0.0 else ds
        dy2 = _PENDING_ADD.pop(dy.data_ptr(), None)         # second addend parked by a lazy ForkFn on this output
        _call("pcm_layernorm_bwd2", dy.data_ptr(), _p(dy2), s.data_ptr(), stat.data_ptr(), gamma.data_ptr(), ds.data_ptr(),
              db.data_ptr(), gg.data_ptr(), gb.data_ptr(), M, E, drop_p, seed, _DT[s.dtype], _s())
        return ds, db, rg, rb, None, None


class MHAFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v per head on the packed in_proj output qkv (B, L, 3E); dropout on the attention
    probabilities as nn.MultiheadAttention does in training."""

    @staticmethod
    def forward(ctx, qkv, n_heads, drop_p, seed):
        qkv = qkv.contiguous()
        B, L, E3 = qkv.shape
        E = E3 // 3
        D = E // n_heads
        out = torch.empty((B, L, E), device=qkv.device, dtype=qkv.dtype)
        lse = torch.empty(B * n_heads * L, device=qkv.device, dtype=torch.float32)
        scale = 1.0 / (D ** 0.5)
        if qkv.dtype == torch.bfloat16 and D == 32 and L <= 224 and os.environ.get("PCM_MHA_TC", "1") != "0":
            _call("pcm_mha_fwd_tc", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, n_heads, scale, drop_p, seed, _s())
        else:
            _call("pcm_mha_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, n_heads, D, scale, drop_p, seed,
                  _DT[qkv.dtype], _s())
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (n_heads, D, scale, drop_p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        n_heads, D, scale, drop_p, seed = ctx.cfg
        B, L, _ = qkv.shape
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        if qkv.dtype == torch.bfloat16 and D == 32 and L <= 224 and os.environ.get("PCM_MHA_TC", "1") != "0":
            _call("pcm_mha_bwd_tc", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, L,
                  n_heads, scale, drop_p, seed, _s())
        else:
            _call("pcm_mha_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, L,
                  n_heads, D, scale, drop_p, seed, _DT[qkv.dtype], _s())
        return dqkv, None, None, None
