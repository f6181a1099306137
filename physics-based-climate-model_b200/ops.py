"""Host-side operator layer: torch.autograd.Functions whose forward/backward bodies are sequences of
pcm_b200 C-ABI kernel launches (include/pcm_b200.h) on torch's current CUDA stream.

Activations travel between Functions as contiguous NHWC tensors (N, H, W, C) in the compute dtype
(torch.bfloat16 by default, torch.float32 for the tight-parity path), C padded to a multiple of 8.
Parameters stay the module's fp32 nn.Parameters in the reference's shapes;
weight gradients are accumulated by the kernels in fp32 either into ``param.main_grad`` (a view of
the trainer's flat, pre-zeroed gradient buffer — no autograd accumulation kernels) or into a fresh
zeroed tensor returned to autograd.

torch is used for device memory (empty/zeros), streams and autograd bookkeeping only; every
arithmetic kernel is ours.  There is no CPU / eager fallback.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import lib

GN_GROUPS = 8
GN_EPS = 1e-5
_DT = {torch.float32: 0, torch.bfloat16: 1}


def pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def padc(c: int) -> int:
    """Channel padding of staged activations: 16 = one UMMA K step of bf16."""
    return (c + 15) // 16 * 16


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(name: str, *args):
    lib().call(name, *args)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"pcm_b200: {what} must be a CUDA tensor — the pcm_b200 kernels have no CPU fallback")


def _grad_buf(p: torch.Tensor):
    """(buffer the kernels accumulate into, value to hand back to autograd)."""
    mg = getattr(p, "main_grad", None)
    if mg is not None:
        return mg, None
    g = torch.zeros_like(p, dtype=torch.float32)
    return g, g


# ------------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# ------------------------------------------------------------------------------------------------
def _work_list(sizes, dev, block: int = 1024) -> torch.Tensor:
    """(job, block) pairs covering every job's elements in blocks of `block` — the grid of the batched kernels."""
    pairs = [(j, b) for j, n in enumerate(sizes) for b in range((n + block - 1) // block)]
    return torch.tensor(pairs, dtype=torch.int32).reshape(-1, 2).to(dev)


class PackPlan:
    """Persistent packed-weight buffers for a training loop: every (parameter, layout) pair the step needs is
    registered the first time it is requested; afterwards `repack()` refreshes ALL of them from the fp32 master
    parameters in ONE launch (pcm_pack_weights_batched) at the start of each step, and `pack_weight` just hands
    out the resident buffer.  Without an active plan (plain module use) weights are packed on demand."""

    def __init__(self):
        self.entries = {}          # key -> (packed tensor, job record)
        self.table = None
        self.dirty = False
        self.max_elems = 0
        # packed weight-gradient buffers ([tap][Co][Cpad] fp32, reduced into with vector atomics by the tensor-core
        # weight-gradient kernel) and the ONE launch that folds them into the parameter-layout gradients
        self.grad_range = None     # (first byte, one-past-last byte) of the optimizer's flat gradient buffer
        self.gentries = {}         # (dst ptr, Co, Ci_tot, taps) -> (packed tensor, job record)
        self.gtables = {}
        self.gdirty = False
        self.gmax = 0
        self.pending_join = False  # repack() was issued on the side stream and nothing has waited for it yet

    def lookup(self, key):
        e = self.entries.get(key)
        return None if e is None else e[0]

    def register(self, key, out, job):
        self.entries[key] = (out, job)
        self.max_elems = max(self.max_elems, out.numel())
        self.dirty = True

    def repack(self):
        if not self.entries:
            return
        if self.dirty:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: a new packed weight was requested after warm-up; run more eager "
                                   "warm-up steps before capturing the graph")
            dev = next(iter(self.entries.values()))[0].device
            self.table = torch.tensor([j for _, j in self.entries.values()], dtype=torch.int64).to(dev)
            self.work = _work_list([t.numel() for t, _ in self.entries.values()], dev)
            self.dirty = False
        _call("pcm_pack_weights_batched", self.table.data_ptr(), self.work.data_ptr(), self.work.shape[0], _s())


    def grad_pack(self, dw: torch.Tensor, Co: int, Ci_tot: int, taps: int):
        """Packed accumulation buffer for the parameter gradient `dw` ((Co, Ci_tot, k, k) fp32 inside the flat
        gradient buffer), or None when dw is not a persistent main_grad view."""
        if self.grad_range is None or not (self.grad_range[0] <= dw.data_ptr() < self.grad_range[1]):
            return None
        key = (dw.data_ptr(), Co, Ci_tot, taps)
        e = self.gentries.get(key)
        if e is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: a new packed gradient was requested after warm-up")
            Cpad = (Ci_tot + 15) // 16 * 16
            if taps * Cpad > 9 * 512:                # slab limit of pcm_unpack_grads_batched: reduce in place instead
                return None
            packed = torch.zeros((taps, Co, Cpad), device=dw.device, dtype=torch.float32)
            job = [packed.data_ptr(), dw.data_ptr(), Ci_tot * taps, taps, 1, Co | (Ci_tot << 32), Cpad | (taps << 32), 0]
            e = (packed, job)
            self.gentries[key] = e
            self.gmax = max(self.gmax, packed.numel())
            self.gdirty = True
        return e[0]

    def unpack_grads(self, lo: Optional[int] = None, hi: Optional[int] = None):
        """Fold the packed buffers into the parameter-layout gradients (one launch).  lo/hi (byte addresses inside
        the flat gradient buffer) restrict the launch to the parameters of one all-reduce bucket."""
        if not self.gentries:
            return
        if self.gdirty:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: packed-gradient table changed during graph capture")
            self.gtables = {}
            self.gdirty = False
        key = (lo, hi)
        if key not in self.gtables:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: packed-gradient bucket requested for the first time during capture")
            sel = [(t, j) for t, j in self.gentries.values() if (lo is None or j[1] >= lo) and (hi is None or j[1] < hi)]
            dev = next(iter(self.gentries.values()))[0].device
            pairs = [(k, co) for k, (t, _) in enumerate(sel) for co in range(t.shape[1])]      # one CTA per (job, co)
            self.gtables[key] = (torch.tensor([j for _, j in sel], dtype=torch.int64).reshape(-1, 8).to(dev),
                                 torch.tensor(pairs, dtype=torch.int32).reshape(-1, 2).to(dev))
        table, work = self.gtables[key]
        if work.shape[0]:
            _call("pcm_unpack_grads_batched", table.data_ptr(), work.data_ptr(), work.shape[0], _s())


_PLAN: Optional[PackPlan] = None
_SIDE: Optional[torch.cuda.Stream] = None      # trainer-provided stream for weight-gradient kernels
_SIDE_FORKED = False                            # work has been issued on _SIDE since the last join


class side_stream:
    """`with side_stream(t1, t2, ...):` issues the enclosed launches on the trainer's side stream (after everything
    enqueued so far on the current stream), so that weight-gradient GEMMs — which nothing in the backward chain
    depends on — overlap the data-gradient / tail kernels of the critical path.  The listed tensors are the ones the
    side kernels read: they are marked in-use on that stream for the caching allocator.  No-op without a side stream."""

    def __init__(self, *tensors):
        self.tensors = tensors
        self.ctx = None

    def __enter__(self):
        global _SIDE_FORKED
        if _SIDE is not None:
            _SIDE_FORKED = True
            _SIDE.wait_stream(torch.cuda.current_stream())
            for t in self.tensors:
                if t is not None:
                    t.record_stream(_SIDE)
            self.ctx = torch.cuda.stream(_SIDE)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


_GRAD_HOOKS = {}          # name -> callable, fired from backward when the gradient at that point has been computed


class GradReadyFn(torch.autograd.Function):
    """Identity whose backward fires the trainer's hook `name`: placed on the ConvLSTM input, it marks the moment in
    backward at which every decoder / head / ConvLSTM weight gradient has been enqueued — the trainer then starts
    the all-reduce of that bucket so that it overlaps the encoder's backward."""

    @staticmethod
    def forward(ctx, x, name):
        ctx.name = name
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        fn = _GRAD_HOOKS.get(ctx.name)
        if fn is not None:
            fn()
        return g, None


def grad_ready_hook(x, name: str):
    return GradReadyFn.apply(x, name) if (x.requires_grad and name in _GRAD_HOOKS) else x


def side_active() -> bool:
    return _SIDE is not None


def side_forked() -> bool:
    """Work has been issued on the side stream since the last join (waiting on it is then legal under graph capture)."""
    return _SIDE is not None and _SIDE_FORKED


def join_side():
    """Make the current stream wait for everything issued on the side stream.  Nothing to do when the side stream
    has not been used since the last join — and under graph capture it MUST be skipped then: waiting on a stream that
    is not part of the capture invalidates it (models without side-stream work, e.g. SimpleCNN)."""
    global _SIDE_FORKED
    if _SIDE is not None and _SIDE_FORKED:
        torch.cuda.current_stream().wait_stream(_SIDE)
        _SIDE_FORKED = False


class use_pack_plan:
    """Context manager activating a PackPlan (and optionally a side stream for weight-gradient kernels) for the ops
    issued inside it (trainer step body)."""

    def __init__(self, plan: Optional[PackPlan], side: Optional[torch.cuda.Stream] = None):
        self.plan = plan
        self.side = side

    def __enter__(self):
        global _PLAN, _SIDE
        self.prev, _PLAN = _PLAN, self.plan
        self.prev_side, _SIDE = _SIDE, self.side
        return self.plan

    def __exit__(self, *exc):
        global _PLAN, _SIDE
        _PLAN = self.prev
        _SIDE = self.prev_side
        return False


def pack_weight(w: torch.Tensor, so: int, si: int, st: int, O: int, I: int, taps: int, dtype, offset: int = 0,
                Op: Optional[int] = None, Ip: Optional[int] = None, group: int = 1):
    """out[t][o][i] = w.flatten()[offset + o*so + i*si + t*st], zero padded to (taps, Op, Ip).
    group > 1 (3x3 kernels only): the pixel-group form (9, group*Op, group*Ip) of pcm_pack_weight_grouped."""
    Op = pad8(O) if Op is None else Op
    Ip = pad8(I) if Ip is None else Ip
    src = w.data_ptr() + 4 * offset
    if _PLAN is not None:
        key = (src, so, si, st, O, I, taps, Op, Ip, dtype, group)
        hit = _PLAN.lookup(key)
        if hit is not None:
            if _PLAN.pending_join:             # the step's re-pack was issued on the side stream: its first consumer waits
                _PLAN.pending_join = False
                join_side()
            return hit
    out = torch.empty((taps, group * Op, group * Ip), device=w.device, dtype=dtype)
    if group > 1:
        assert taps == 9
        _call("pcm_pack_weight_grouped", src, so, si, st, O, I, Op, Ip, group, out.data_ptr(), _DT[dtype], _s())
    else:
        _call("pcm_pack_weight", src, so, si, st, O, I, taps, Op, Ip, out.data_ptr(), _DT[dtype], _s())
    if _PLAN is not None:
        _PLAN.register(key, out, [src, out.data_ptr(), so, si, st, O | (I << 32), taps | (Op << 32),
                                  Ip | ((_DT[dtype] | ((group if group > 1 else 0) << 8)) << 32)])
    return out


def conv_group(dtype, Sc: int, Dc: int, W: int, dense: bool = True) -> int:
    """Pixel-group factor for a thin 3x3 layer on the tensor-core path (pcm_conv3x3_tc_grouped): g adjacent pixels
    of a row become one GEMM row of g*Sc channels, which divides the number of TMA box rows — the measured bound of
    the 16- and 32-channel layers — by g.  1 = plain form.  PCM_CONV_GROUP=0 disables, =2 caps the factor at 2;
    PCM_CONV_GROUP_MAXC (default 32) is the widest layer (max of Sc, Dc) that is grouped (32-channel layers in pairs:
    -10 us per step)."""
    import os
    cap = int(os.environ.get("PCM_CONV_GROUP", "4"))
    maxc = int(os.environ.get("PCM_CONV_GROUP_MAXC", "32"))
    if cap < 2 or not dense or dtype != torch.bfloat16 or Sc % 16 or Dc % 16 or max(Sc, Dc) > min(maxc, 32):
        return 1
    g = min(cap, 64 // max(Sc, Dc))
    g = 4 if g >= 4 else 2 if g >= 2 else 1
    while g > 1 and W % g:
        g //= 2
    return g


def conv_weight_fwd(w: torch.Tensor, dtype, ci_off: int = 0, ci: Optional[int] = None, Ip: Optional[int] = None,
                    co_split: int = 0, group: int = 1):
    """nn.Conv2d weight (Co, Ci_tot, KH, KW) -> [taps][Co][Ci] (forward, gather mode 0).  co_split > 0: a list of
    packed slices of co_split output channels each (see conv_s1).  group > 1: pixel-group form (conv_group)."""
    Co, Ci_tot, KH, KW = w.shape
    ci = Ci_tot - ci_off if ci is None else ci
    K = KH * KW
    if co_split and Co > co_split:
        assert Co % co_split == 0
        return [pack_weight(w, Ci_tot * K, K, 1, co_split, ci, K, dtype, offset=c0 * Ci_tot * K + ci_off * K, Ip=Ip)
                for c0 in range(0, Co, co_split)]
    return pack_weight(w, Ci_tot * K, K, 1, Co, ci, K, dtype, offset=ci_off * K, Ip=Ip, group=group)


def conv_weight_dgrad(w: torch.Tensor, dtype, ci_off: int = 0, ci: Optional[int] = None, Op: Optional[int] = None,
                      group: int = 1):
    """nn.Conv2d weight (stride 1, odd K) -> [taps][Ci][Co] with the taps FLIPPED, so the data gradient
    is itself a forward (mode 0) convolution of dy: dx = conv(dy, flip(W)^T).  group > 1: pixel-group form."""
    Co, Ci_tot, KH, KW = w.shape
    ci = Ci_tot - ci_off if ci is None else ci
    K = KH * KW
    return pack_weight(w, K, Ci_tot * K, -1, ci, Co, K, dtype, offset=ci_off * K + K - 1, Op=Op, group=group)


def tc_supported(dtype, Sc: int, Dc: int, K: int = 3) -> bool:
    """Shapes the tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu) takes; everything else runs on the
    general-shape SIMT gather kernel."""
    return (dtype == torch.bfloat16 and K == 3 and (Sc in (16, 32) or (Sc >= 64 and Sc % 64 == 0))
            and Dc % 16 == 0 and 16 <= Dc <= 256)


def conv_s1(src, wk, N, H, W, Sc, Dc, K=3, dst=None, dst_f32=False, bias=None, accumulate=False,
            src_ns=None, src_ps=None, dst_ns=None, dst_ps=None, src_off=0, dst_off=0, group=1):
    """Stride-1 'same' convolution dst = conv(src, wk) on NHWC views; wk = [K*K][Dc][Sc], or a list of such tensors
    holding consecutive output-channel slices (outputs wider than one 256-column accumulator, e.g. the 512 gate
    channels of the base-32 ConvLSTM: one launch per slice into its channel range of dst).
    group > 1: wk is the pixel-group packing [9][group*Dc][group*Sc] (conv_group / conv_weight_*(group=...)); dense
    pixels on both sides, no bias."""
    dtype = src.dtype
    if group > 1:
        assert K == 3 and bias is None and not accumulate and src_ps in (None, Sc) and dst_ps in (None, Dc)
        if dst is None:
            dst = torch.empty((N, H, W, Dc), device=src.device, dtype=torch.float32 if dst_f32 else dtype)
        _call("pcm_conv3x3_tc_grouped", src.data_ptr() + src_off * src.element_size(),
              H * W * Sc if src_ns is None else src_ns, Sc, H, W, Sc, dst.data_ptr() + dst_off * dst.element_size(),
              H * W * Dc if dst_ns is None else dst_ns, Dc, Dc, wk.data_ptr(), N, int(dst_f32), group, _s())
        return dst
    if isinstance(wk, (list, tuple)):
        if dst is None:
            dst = torch.empty((N, H, W, Dc), device=src.device, dtype=torch.float32 if dst_f32 else dtype)
        dps = Dc if dst_ps is None else dst_ps
        c0 = 0
        for part in wk:
            dc = part.shape[1]
            conv_s1(src, part, N, H, W, Sc, dc, K, dst=dst, dst_f32=dst_f32, bias=None if bias is None else bias[c0:c0 + dc],
                    accumulate=accumulate, src_ns=src_ns, src_ps=src_ps, dst_ns=H * W * dps if dst_ns is None else dst_ns,
                    dst_ps=dps, src_off=src_off, dst_off=dst_off + c0)
            c0 += dc
        return dst
    if not tc_supported(dtype, Sc, Dc, K):
        return conv_gather(src, wk, N, H, W, Sc, H, W, Dc, K, K, 1, K // 2, 0, dst=dst, dst_f32=dst_f32, bias=bias,
                           accumulate=accumulate, src_ns=src_ns, src_ps=src_ps, dst_ns=dst_ns, dst_ps=dst_ps,
                           src_off=src_off, dst_off=dst_off)
    if dst is None:
        dst = torch.empty((N, H, W, Dc), device=src.device, dtype=torch.float32 if dst_f32 else dtype)
    src_ps = Sc if src_ps is None else src_ps
    dst_ps = Dc if dst_ps is None else dst_ps
    src_ns = H * W * src_ps if src_ns is None else src_ns
    dst_ns = H * W * dst_ps if dst_ns is None else dst_ns
    _call("pcm_conv3x3_tc", src.data_ptr() + src_off * src.element_size(), src_ns, src_ps, H, W, Sc,
          dst.data_ptr() + dst_off * dst.element_size(), dst_ns, dst_ps, Dc, wk.data_ptr(), _p(bias), N,
          int(dst_f32), int(accumulate), _s())
    return dst


def conv_gather(src, wk, N, Hs, Ws, Sc, Hd, Wd, Dc, KH, KW, stride, pad, mode, dst=None, dst_f32=False,
                bias=None, accumulate=False, relu=False, src_ns=None, src_ps=None, dst_ns=None, dst_ps=None,
                src_off=0, dst_off=0, dtype=None):
    dtype = dtype or src.dtype
    if dst is None:
        dst = torch.empty((N, Hd, Wd, Dc), device=src.device, dtype=torch.float32 if dst_f32 else dtype)
    src_ps = Sc if src_ps is None else src_ps
    dst_ps = Dc if dst_ps is None else dst_ps
    src_ns = Hs * Ws * src_ps if src_ns is None else src_ns
    dst_ns = Hd * Wd * dst_ps if dst_ns is None else dst_ns
    _call("pcm_conv_gather", src.data_ptr() + src_off * src.element_size(), src_ns, src_ps, Hs, Ws, Sc,
          dst.data_ptr() + dst_off * dst.element_size(), dst_ns, dst_ps, Hd, Wd, Dc, wk.data_ptr(), _p(bias),
          N, KH, KW, stride, pad, mode, int(dst_f32), int(accumulate), int(relu), _DT[dtype], _s())
    return dst


def conv_wgrad(A, B, dw, sa, sb, st, N, Ha, Wa, Ca, Ca_real, Hb, Wb, Cb, Cb_real, KH, KW, stride, pad,
               a_ns=None, a_ps=None, b_ns=None, b_ps=None, a_off=0, b_off=0, dw_off=0):
    a_ps = Ca if a_ps is None else a_ps
    b_ps = Cb if b_ps is None else b_ps
    a_ns = Ha * Wa * a_ps if a_ns is None else a_ns
    b_ns = Hb * Wb * b_ps if b_ns is None else b_ns
    _call("pcm_conv_wgrad", A.data_ptr() + a_off * A.element_size(), a_ns, a_ps, Ha, Wa, Ca, Ca_real,
          B.data_ptr() + b_off * B.element_size(), b_ns, b_ps, Hb, Wb, Cb, Cb_real,
          dw.data_ptr() + 4 * dw_off, sa, sb, st, N, KH, KW, stride, pad, _DT[A.dtype], _s())


def wgrad_tc_supported(dtype, Co: int, Ci: int, H: int, W: int) -> bool:
    # rows wider than one TMA box (W + 2 > 256, config 5) are covered in column strips inside pcm_wgrad3x3_tc
    return (dtype == torch.bfloat16 and (Co in (16, 32, 64) or (Co % 128 == 0 and Co > 0))
            and Ci in (16, 32, 64, 128, 192, 256) and H + 2 <= 256)


import os as _os
# opt-in (measured neutral: the stand-alone kernel costs the side stream more than the tail saves, profiles/r2_experiments.md)
_GATE_WGRAD_SIDE = _os.environ.get("PCM_GATE_WGRAD_SIDE", "0") == "1"

# The backward of MaxPool2d (+ time-mean skip) holds the gradient it produces and the block output it routes through in
# registers, and the ConvBlock backward that runs NEXT needs exactly their per-pixel product sum (the gate gradient):
# the pooling kernel writes it (pcm_maxpool2_bwd_skip_dot) and hands it over through this one-entry slot.  The slot
# keeps the gradient tensor alive (its address cannot be reused for something else meanwhile) and is emptied by the
# next ConvBlock backward whatever it finds.  PCM_POOL_SDOT=0: the tail streams dout and out itself.
_SDOT_SLOT = None          # (gradient tensor, sdot tensor)


def _pool_sdot_enabled() -> bool:
    return _os.environ.get("PCM_POOL_SDOT", "1") != "0"


def _offer_sdot(ds: torch.Tensor):
    """Called by the pooling backward before its launch: the (N*H*W,) fp32 buffer to fill, or None."""
    global _SDOT_SLOT
    _SDOT_SLOT = None
    N, H, W, C = ds.shape
    cv = C // 8
    if not _pool_sdot_enabled() or C % 8 or cv > 32 or cv & (cv - 1) or ds.dtype not in _DT:
        return None
    # only the shared-memory resident tail takes the sum (SEBlock ratio 8: Cr = C // 8); images that do not fit one SM
    # (config 5) go through the multi-kernel path, which forms dq itself
    if ds.is_cuda and not lib()._fn["pcm_convblock_fused_supported"](H, W, C, max(C // 8, 1), _DT[ds.dtype]):
        return None
    sdot = torch.empty(N * H * W, device=ds.device, dtype=torch.float32)
    _SDOT_SLOT = (ds, sdot)
    return sdot


def _take_sdot(dout: torch.Tensor):
    """Called by every ConvBlock backward: the sum for exactly this gradient tensor, or None.  Empties the slot."""
    global _SDOT_SLOT
    slot, _SDOT_SLOT = _SDOT_SLOT, None
    if slot is None:
        return None
    ds, sdot = slot
    if ds.data_ptr() == dout.data_ptr() and ds.shape == dout.shape and ds.dtype == dout.dtype and dout.is_contiguous():
        return sdot
    return None


def wgrad_group(dtype, Co: int, Ci: int, W: int, dense: bool = True) -> int:
    """Pixel-group factor of the tensor-core weight gradient (pcm_wgrad3x3_tc_grouped) for the thin layers; 1 = plain.
    PCM_WGRAD_GROUP=0 disables, =2 caps the factor; PCM_WGRAD_GROUP_MAXC (default 32) = widest layer that is grouped."""
    import os
    cap = int(os.environ.get("PCM_WGRAD_GROUP", "4"))
    maxc = int(os.environ.get("PCM_WGRAD_GROUP_MAXC", "32"))
    if cap < 2 or not dense or dtype != torch.bfloat16 or Co not in (16, 32) or Ci not in (16, 32) or max(Co, Ci) > maxc:
        return 1
    g = min(cap, 64 // max(Co, Ci))
    g = 4 if g >= 4 else 2 if g >= 2 else 1
    while g > 1 and (W % g or 3 * (g + 2) * Ci > 512):
        g //= 2
    return g


def conv3x3_wgrad(dy, x, dw, N, H, W, Co, Ci, Ci_real, Ci_tot=None, dy_ns=None, dy_ps=None, x_ns=None, x_ps=None,
                  dy_off=0, x_off=0, dw_off=0):
    """dw[co][ci_off + ci][tap] += sum_p dy(p, co) * x(p + tap, ci) for the 3x3/s1/p1 convolution.
    dw is the fp32 master-layout gradient (Co, Ci_tot, 3, 3); dw_off selects a ci slice (ConvLSTM Wx / Wh)."""
    Ci_tot = Ci_real if Ci_tot is None else Ci_tot
    dy_ps = Co if dy_ps is None else dy_ps
    x_ps = Ci if x_ps is None else x_ps
    dy_ns = H * W * dy_ps if dy_ns is None else dy_ns
    x_ns = H * W * x_ps if x_ns is None else x_ns
    if wgrad_tc_supported(dy.dtype, Co, Ci, H, W):
        packed = _PLAN.grad_pack(dw, Co, Ci_tot, 9) if _PLAN is not None else None
        grp = wgrad_group(dy.dtype, Co, Ci, W, dy_ps == Co and x_ps == Ci)
        if grp > 1:
            if packed is not None:
                dst, sa, sb, st = packed.data_ptr() + 4 * (dw_off // 9), packed.shape[-1], 1, Co * packed.shape[-1]
            else:
                dst, sa, sb, st = dw.data_ptr() + 4 * dw_off, Ci_tot * 9, 9, 1
            _call("pcm_wgrad3x3_tc_grouped", dy.data_ptr() + dy_off * dy.element_size(), dy_ns, Co,
                  x.data_ptr() + x_off * x.element_size(), x_ns, Ci, Ci_real, dst, sa, sb, st, N, H, W, grp, _s())
            return
        if packed is not None:
            # reduce into the packed [tap][Co][Cpad] buffer (16-byte vector atomics); PackPlan.unpack_grads folds
            # every layer's buffer into the parameter-layout gradient in one launch after backward
            Cpad = packed.shape[-1]
            _call("pcm_wgrad3x3_tc", dy.data_ptr() + dy_off * dy.element_size(), dy_ns, dy_ps, Co, Co,
                  x.data_ptr() + x_off * x.element_size(), x_ns, x_ps, Ci, Ci_real, packed.data_ptr() + 4 * (dw_off // 9),
                  Cpad, 1, Co * Cpad, N, H, W, _s())
            return
        _call("pcm_wgrad3x3_tc", dy.data_ptr() + dy_off * dy.element_size(), dy_ns, dy_ps, Co, Co,
              x.data_ptr() + x_off * x.element_size(), x_ns, x_ps, Ci, Ci_real, dw.data_ptr() + 4 * dw_off,
              Ci_tot * 9, 9, 1, N, H, W, _s())
    else:
        conv_wgrad(dy, x, dw, Ci_tot * 9, 9, 1, N, H, W, Co, Co, H, W, Ci, Ci_real, 3, 3, 1, 1, a_ns=dy_ns, a_ps=dy_ps,
                   b_ns=x_ns, b_ps=x_ps, a_off=dy_off, b_off=x_off, dw_off=dw_off)


# ---- ConvTranspose2d(kernel 2, stride 2): tensor-core GEMM forms, SIMT gather kernels otherwise ----------------
def _tc_k_ok(c: int) -> bool:
    return c in (16, 32) or (c >= 64 and c % 64 == 0)


def convT2x2_fwd(x, wt, bt, B, h, w, Ci, Co, dst, dst_ns, dst_ps, relu=False):
    """dst(b, 2h+kh, 2w+kw, co) = sum_ci x(b,h,w,ci) * wt[ci][co][kh][kw] + bt[co]; dst may be a channel slice of a
    wider buffer (dst_ns / dst_ps = its image / pixel strides)."""
    dt = x.dtype
    if dt == torch.bfloat16 and _tc_k_ok(Ci) and Co > 64 and Co % 64 == 0:
        # the GEMM's N extent is 4*Co <= 256: wider outputs go in slices of 64 channels of the destination
        for c0 in range(0, Co, 64):
            wk = pack_weight(wt, 4, Co * 4, 1, 64, Ci, 4, dt, offset=c0 * 4)
            _call("pcm_convT2x2_tc", x.data_ptr(), h * w * Ci, Ci, h, w, Ci, dst.data_ptr() + c0 * dst.element_size(), dst_ns,
                  dst_ps, 64, wk.data_ptr(), 0 if bt is None else bt.data_ptr() + 4 * c0, B, int(relu), _s())
        return dst
    wk = pack_weight(wt, 4, Co * 4, 1, Co, Ci, 4, dt)                 # wk[tap][co][ci] = wt[ci][co][tap]
    if dt == torch.bfloat16 and _tc_k_ok(Ci) and Co % 16 == 0 and Co <= 64:
        _call("pcm_convT2x2_tc", x.data_ptr(), h * w * Ci, Ci, h, w, Ci, dst.data_ptr(), dst_ns, dst_ps, Co, wk.data_ptr(),
              _p(bt), B, int(relu), _s())
    else:
        conv_gather(x, wk, B, h, w, Ci, 2 * h, 2 * w, Co, 2, 2, 2, 0, 1, dst=dst, bias=bt, relu=relu, dst_ns=dst_ns,
                    dst_ps=dst_ps)
    return dst


def convT2x2_dgrad(dy, wt, B, h, w, Ci, Co, dy_ns, dy_ps):
    """dx(b,h,w,ci) = sum_{kh,kw,co} dy(b, 2h+kh, 2w+kw, co) * wt[ci][co][kh][kw]."""
    dt = dy.dtype
    wk = pack_weight(wt, Co * 4, 4, 1, Ci, Co, 4, dt)                 # wk[tap][ci][co] = wt[ci][co][tap]
    dx = torch.empty((B, h, w, Ci), device=dy.device, dtype=dt)
    if dt == torch.bfloat16 and _tc_k_ok(Co) and Ci % 16 == 0 and Ci <= 256:
        _call("pcm_convT2x2_dgrad_tc", dy.data_ptr(), dy_ns, dy_ps, h, w, Co, dx.data_ptr(), h * w * Ci, Ci, Ci,
              wk.data_ptr(), B, _s())
    else:
        conv_gather(dy, wk, B, 2 * h, 2 * w, Co, h, w, Ci, 2, 2, 2, 0, 0, dst=dx, src_ns=dy_ns, src_ps=dy_ps)
    return dx


def convT2x2_wgrad(x, dy, gwt, B, h, w, Ci, Co, dy_ns, dy_ps):
    """gwt (Ci, Co, 2, 2) fp32 += sum_{b,h,w} x(b,h,w,ci) * dy(b, 2h+kh, 2w+kw, co)."""
    if (x.dtype == torch.bfloat16 and (Ci in (16, 32, 64) or Ci % 128 == 0) and Co in (16, 32, 64, 128, 192, 256)
            and w <= 256 and h <= 256):
        packed = _PLAN.grad_pack(gwt, Ci, Co, 4) if _PLAN is not None else None
        if packed is not None:
            Cpad = packed.shape[-1]
            _call("pcm_convT2x2_wgrad_tc", x.data_ptr(), h * w * Ci, Ci, Ci, Ci, dy.data_ptr(), dy_ns, dy_ps, Co, Co,
                  packed.data_ptr(), Cpad, 1, Ci * Cpad, B, h, w, _s())
        else:
            _call("pcm_convT2x2_wgrad_tc", x.data_ptr(), h * w * Ci, Ci, Ci, Ci, dy.data_ptr(), dy_ns, dy_ps, Co, Co,
                  gwt.data_ptr(), Co * 4, 4, 1, B, h, w, _s())
    else:
        conv_wgrad(x, dy, gwt, Co * 4, 4, 1, B, h, w, Ci, Ci, 2 * h, 2 * w, Co, Co, 2, 2, 2, 0, b_ns=dy_ns, b_ps=dy_ps)


def channel_sum(x, out, N, P, C, C_real, ns=None, ps=None, off=0, per_image=False):
    ps = C if ps is None else ps
    ns = P * ps if ns is None else ns
    _call("pcm_channel_sum", x.data_ptr() + off * x.element_size(), ns, ps, N, P, C, C_real, out.data_ptr(),
          int(per_image), _DT[x.dtype], _s())


# ------------------------------------------------------------------------------------------------
# layout staging at the module boundary (NCHW fp32 <-> NHWC compute dtype)
# ------------------------------------------------------------------------------------------------
class StageIn(torch.autograd.Function):
    """(N, C, H, W) fp32 -> (N, H, W, padc(C)) compute dtype."""

    @staticmethod
    def forward(ctx, x, dtype, gran=16, T=1):
        """T > 1: x is a flattened (B, T) window batch (image b*T + t); the staged tensor is t-major."""
        _require_cuda(x, "input")
        x = x.contiguous().float()
        N, C, H, W = x.shape
        Cp = (C + gran - 1) // gran * gran
        y = torch.empty((N, H, W, Cp), device=x.device, dtype=dtype)
        _call("pcm_nchw_to_nhwc", x.data_ptr(), y.data_ptr(), N, C, H, W, Cp, T, _DT[dtype], _s())
        ctx.shape = (N, C, H, W, T)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W, T = ctx.shape
        dy = dy.contiguous()
        dx = torch.empty((N, C, H, W), device=dy.device, dtype=torch.float32)
        _call("pcm_nhwc_to_nchw", dy.data_ptr(), dx.data_ptr(), N, C, H, W, dy.shape[-1], T, _DT[dy.dtype], _s())
        return dx, None, None, None


class StageOut(torch.autograd.Function):
    """(N, H, W, Cp) compute dtype -> (N, C, H, W) fp32."""

    @staticmethod
    def forward(ctx, x, C):
        x = x.contiguous()
        N, H, W, Cp = x.shape
        y = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
        _call("pcm_nhwc_to_nchw", x.data_ptr(), y.data_ptr(), N, C, H, W, Cp, 1, _DT[x.dtype], _s())
        ctx.meta = (Cp, x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        Cp, dtype = ctx.meta
        dy = dy.contiguous().float()
        N, C, H, W = dy.shape
        dx = torch.empty((N, H, W, Cp), device=dy.device, dtype=dtype)
        _call("pcm_nchw_to_nhwc", dy.data_ptr(), dx.data_ptr(), N, C, H, W, Cp, 1, _DT[dtype], _s())
        return dx, None


def season_embed_stage(x5: torch.Tensor, month: torch.Tensor, dtype, T: int = 1) -> torch.Tensor:
    """main_final.py:186-216: (N,5,H,W) forcings + month index (N,) -> NHWC 8-channel frames with
    sin/cos month channels (5, 6) synthesised on the fly, padded to 16 channels (no gradient: inputs are data).
    T > 1: the N = B*T frames are given window-major (b*T + t) and staged t-major, as AttUNetConvLSTM wants."""
    _require_cuda(x5, "input")
    N, C, H, W = x5.shape
    assert C == 5
    y = torch.empty((N, H, W, 16), device=x5.device, dtype=dtype)
    _call("pcm_season_embed_stage", x5.contiguous().float().data_ptr(), month.to(torch.int32).contiguous().data_ptr(),
          y.data_ptr(), N, H, W, 16, T, _DT[dtype], _s())
    return y


def window_stage(series: torch.Tensor, idx: torch.Tensor, T: int, dtype, norm: Optional[torch.Tensor] = None,
                 month: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SequenceDataset.__getitem__ (main_final.py:97-154) for a whole batch, on the device: `series` is the resident
    input record (Ttot, C, H, W) fp32, `idx` (B,) the target time indices; returns the t-major staged frames
    (T*B, H, W, 16): frame (t, b) = series[idx[b] - T + 1 + t], zeros where that index is negative (left pad).
    norm (C, 4) fp64 [data.Normalizer.input_table]: the record is RAW and Normalizer.normalize is applied while staging;
    month (Ttot,) int32: sin/cos month channels C, C+1 are synthesised (main_final.py:186-216).
    Inputs are data: no gradient."""
    _require_cuda(series, "series")
    Ttot, C, H, W = series.shape
    B = idx.numel()
    Cout = C + (2 if month is not None else 0)
    # frame table: a (T, B) int tensor computed with torch integer ops (index plumbing, no activation arithmetic)
    frames = (idx.to(torch.int32).reshape(1, B) - (T - 1) + torch.arange(T, device=idx.device, dtype=torch.int32).reshape(T, 1))
    frames = torch.where(frames < Ttot, frames, torch.full_like(frames, -1)).contiguous()
    y = torch.empty((T * B, H, W, padc(Cout)), device=series.device, dtype=dtype)
    if norm is not None:
        assert norm.dtype == torch.float64 and tuple(norm.shape) == (C, 4) and norm.is_cuda and norm.is_contiguous()
    if month is not None:
        assert month.dtype == torch.int32 and month.numel() == Ttot and month.is_cuda and month.is_contiguous()
    _call("pcm_window_stage", series.contiguous().float().data_ptr(), frames.data_ptr(), y.data_ptr(), T * B, C, H, W,
          padc(Cout), _p(norm), _p(month), _DT[dtype], _s())
    return y


# ------------------------------------------------------------------------------------------------
# ConvBlock: [conv3x3 -> GN(8) -> SiLU] x2 -> SE -> SpatialGate   (src/unet.py:32-49)
# ------------------------------------------------------------------------------------------------
def convblock_fused_ok(H: int, W: int, C: int, Cr: int, dtype) -> bool:
    """True when one H x W x C image (+ gate maps) fits an SM's shared memory: the per-image fused tails
    (csrc/convblock_fused.cu) then replace the grid-wide multi-kernel tails.  PCM_FUSED_TAILS=0 disables them."""
    import os
    if os.environ.get("PCM_FUSED_TAILS", "1") == "0":
        return False
    return bool(lib()._fn["pcm_convblock_fused_supported"](H, W, C, Cr, _DT[dtype]))


def convblock_fwd_tc_ok(H: int, W: int, Cin: int, C: int, Cr: int, dtype) -> bool:
    """True when the whole-block forward kernel (csrc/convblock_fused.cu, convblock_fwd_tc_kernel) is selected for this
    shape.  It is OPT-IN (PCM_BLOCK_FWD_TC=1): parity-green, but measured on B200 it only ties the 4-kernel forward at
    48x72x16 (178-188 us vs 188 us for 384 images) and loses at the smaller levels, where the 4-kernel path overlaps
    several small CTAs per SM (profiles/r2_convblock_fwd_tc.md has the per-phase cycle counts)."""
    import os
    if dtype != torch.bfloat16 or os.environ.get("PCM_BLOCK_FWD_TC", "0") != "1":
        return False
    return bool(lib()._fn["pcm_convblock_fwd_tc_supported"](H, W, Cin, C, Cr))


class ConvBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp):
        _require_cuda(x, "activation")
        x = x.contiguous()
        N, H, W, Cip = x.shape
        Co, Ci = w1.shape[0], w1.shape[1]
        assert Ci <= Cip and Cip % 8 == 0 and Co % 8 == 0, (Ci, Cip, Co)
        Cr = sw1.shape[0]
        P, G, dt, dev = H * W, GN_GROUPS, x.dtype, x.device
        d = _DT[dt]
        st = _s()
        fused = convblock_fused_ok(H, W, Co, Cr, dt)
        ctx.fused = fused
        ctx.dims = (N, H, W, Cip, Ci, Co, Cr)
        ga, gb = (conv_group(dt, Cip, Co, W), conv_group(dt, Co, Co, W)) if fused else (1, 1)
        if fused and convblock_fwd_tc_ok(H, W, Cip, Co, Cr, dt):
            ga = gb = 1
        wk1 = conv_weight_fwd(w1, dt, Ip=Cip, group=ga)
        wk2 = conv_weight_fwd(w2, dt, group=gb)
        if fused and convblock_fwd_tc_ok(H, W, Cip, Co, Cr, dt):
            # the whole block forward in ONE launch: both convolutions on the tensor cores over the shared-memory
            # resident image, conv outputs in tensor memory, a1 written straight into conv2's operand image
            small = torch.empty(N * G * 2 * 2 + N * Co, device=dev, dtype=torch.float32)
            stats1, stats2, pool = small[: N * G * 2], small[N * G * 2: N * G * 4], small[N * G * 4:]
            se = torch.empty(N * Co + N * Cr, device=dev, dtype=torch.float32)
            hid = se[N * Co:]
            y1 = torch.empty((N, H, W, Co), device=dev, dtype=dt)
            a1, y2, out = torch.empty_like(y1), torch.empty_like(y1), torch.empty_like(y1)
            maps = torch.empty(N * P * 3, device=dev, dtype=torch.float32)
            ties = torch.empty(N * P, device=dev, dtype=torch.uint8)
            _call("pcm_convblock_fwd_tc", x.data_ptr(), wk1.data_ptr(), wk2.data_ptr(), g1.data_ptr(), b1.data_ptr(),
                  g2.data_ptr(), b2.data_ptr(), sw1.data_ptr(), sw2.data_ptr(), wsp.data_ptr(), y1.data_ptr(), a1.data_ptr(),
                  y2.data_ptr(), stats1.data_ptr(), stats2.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(),
                  maps.data_ptr(), ties.data_ptr(), out.data_ptr(), N, H, W, Cip, Co, Cr, GN_EPS, st)
            ctx.save_for_backward(x, y1, a1, y2, small, se, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp, maps, ties, out)
            return out
        if fused:
            # per-image tails: conv -> [GN+SiLU] -> conv -> [GN+SiLU+SE+gate]; 4 launches, every tensor read once
            small = torch.empty(N * G * 2 * 2 + N * Co, device=dev, dtype=torch.float32)     # written fully
            stats1, stats2, pool = small[: N * G * 2], small[N * G * 2: N * G * 4], small[N * G * 4:]
            se = torch.empty(N * Co + N * Cr, device=dev, dtype=torch.float32)
            hid = se[N * Co:]
            y1 = conv_s1(x, wk1, N, H, W, Cip, Co, group=ga)
            a1 = torch.empty_like(y1)
            _call("pcm_gn_silu_img_fwd", y1.data_ptr(), g1.data_ptr(), b1.data_ptr(), stats1.data_ptr(), a1.data_ptr(),
                  N, H, W, Co, GN_EPS, d, st)
            y2 = conv_s1(a1, wk2, N, H, W, Co, Co, group=gb)
            out = torch.empty_like(y2)
            # saved for the backward tail: channel mean / max maps and the gate (fp32) + tie counts of the max
            maps = torch.empty(N * P * 3, device=dev, dtype=torch.float32)
            ties = torch.empty(N * P, device=dev, dtype=torch.uint8)
            _call("pcm_convblock_tail_fwd", y2.data_ptr(), g2.data_ptr(), b2.data_ptr(), sw1.data_ptr(), sw2.data_ptr(),
                  wsp.data_ptr(), stats2.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(), maps.data_ptr(),
                  ties.data_ptr(), out.data_ptr(), N, H, W, Co, Cr, GN_EPS, d, st)
            ctx.save_for_backward(x, y1, a1, y2, small, se, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp, maps, ties, out)
            return out
        # one zeroed scratch for every small accumulator of the forward
        small = torch.zeros(N * G * 2 * 2 + N * Co, device=dev, dtype=torch.float32)
        stats1 = small[: N * G * 2]
        stats2 = small[N * G * 2: N * G * 4]
        pool = small[N * G * 4:]
        y1 = conv_s1(x, wk1, N, H, W, Cip, Co)
        _call("pcm_gn_stats", y1.data_ptr(), stats1.data_ptr(), N, P, Co, G, d, st)
        a1 = torch.empty_like(y1)
        _call("pcm_gn_silu_fwd", y1.data_ptr(), stats1.data_ptr(), g1.data_ptr(), b1.data_ptr(), a1.data_ptr(), 0,
              N, P, Co, G, GN_EPS, d, st)
        y2 = conv_s1(a1, wk2, N, H, W, Co, Co)
        _call("pcm_gn_stats", y2.data_ptr(), stats2.data_ptr(), N, P, Co, G, d, st)
        a2 = torch.empty_like(y2)
        _call("pcm_gn_silu_fwd", y2.data_ptr(), stats2.data_ptr(), g2.data_ptr(), b2.data_ptr(), a2.data_ptr(),
              pool.data_ptr(), N, P, Co, G, GN_EPS, d, st)
        se = torch.empty(N * Co + N * Cr, device=dev, dtype=torch.float32)
        hid = se[N * Co:]
        maps = torch.empty(N * P * 3, device=dev, dtype=torch.float32)
        cmap, gate = maps[: N * P * 2], maps[N * P * 2:]
        _call("pcm_se_chanstat_fwd", a2.data_ptr(), pool.data_ptr(), sw1.data_ptr(), sw2.data_ptr(), se.data_ptr(),
              hid.data_ptr(), cmap.data_ptr(), N, P, Co, Cr, d, st)
        out = torch.empty_like(a2)
        _call("pcm_spatial_gate_fwd", a2.data_ptr(), se.data_ptr(), cmap.data_ptr(), wsp.data_ptr(), gate.data_ptr(),
              out.data_ptr(), N, H, W, Co, d, st)
        ctx.save_for_backward(x, y1, a1, y2, a2, small, se, maps, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp)
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.fused:
            return ConvBlockFn._backward_fused(ctx, dout)
        _take_sdot(dout)                                # not used on this path; the slot must not outlive this backward
        x, y1, a1, y2, a2, small, se, maps, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp = ctx.saved_tensors
        N, H, W, Cip, Ci, Co, Cr = ctx.dims
        P, G, dt, dev = H * W, GN_GROUPS, x.dtype, x.device
        d = _DT[dt]
        st = _s()
        dout = dout.contiguous()
        stats1 = small[: N * G * 2]
        stats2 = small[N * G * 2: N * G * 4]
        pool = small[N * G * 4:]
        hid = se[N * Co:]
        cmap, gate = maps[: N * P * 2], maps[N * P * 2:]
        # scratch: dse[N*Co] | gsum2[N*G*2] | gsum1[N*G*2]   (zeroed) ; dq, dpool (written fully)
        z = torch.zeros(N * Co + N * G * 4, device=dev, dtype=torch.float32)
        dse, gsum2, gsum1 = z[: N * Co], z[N * Co: N * Co + N * G * 2], z[N * Co + N * G * 2:]
        e = torch.empty(N * P + N * Co, device=dev, dtype=torch.float32)
        dq, dpool = e[: N * P], e[N * P:]
        gw1, rw1 = _grad_buf(w1); gg1, rg1 = _grad_buf(g1); gb1, rb1 = _grad_buf(b1)
        gw2, rw2 = _grad_buf(w2); gg2, rg2 = _grad_buf(g2); gb2, rb2 = _grad_buf(b2)
        gs1, rs1 = _grad_buf(sw1); gs2, rs2 = _grad_buf(sw2); gsp, rsp = _grad_buf(wsp)

        _call("pcm_spatial_gate_bwd_dq", dout.data_ptr(), a2.data_ptr(), se.data_ptr(), gate.data_ptr(), dq.data_ptr(),
              N, P, Co, d, st)
        with side_stream(e, maps):              # a parameter gradient: nothing in the backward chain waits for it
            _call("pcm_spatial_gate_bwd_dw", dq.data_ptr(), cmap.data_ptr(), gsp.data_ptr(), N, H, W, _s())
        da2 = torch.empty_like(a2)
        _call("pcm_spatial_gate_bwd_da", dout.data_ptr(), a2.data_ptr(), se.data_ptr(), gate.data_ptr(),
              cmap.data_ptr(), dq.data_ptr(), wsp.data_ptr(), da2.data_ptr(), dse.data_ptr(), N, H, W, Co, d, st)
        _call("pcm_se_bwd", dse.data_ptr(), se.data_ptr(), hid.data_ptr(), pool.data_ptr(), sw1.data_ptr(),
              sw2.data_ptr(), dpool.data_ptr(), gs1.data_ptr(), gs2.data_ptr(), N, P, Co, Cr, st)
        _call("pcm_gn_silu_bwd_reduce", da2.data_ptr(), dpool.data_ptr(), y2.data_ptr(), stats2.data_ptr(),
              g2.data_ptr(), b2.data_ptr(), gsum2.data_ptr(), gg2.data_ptr(), gb2.data_ptr(), N, P, Co, G, GN_EPS, d, st)
        dy2 = torch.empty_like(y2)
        _call("pcm_gn_silu_bwd_apply", da2.data_ptr(), dpool.data_ptr(), y2.data_ptr(), stats2.data_ptr(),
              g2.data_ptr(), b2.data_ptr(), gsum2.data_ptr(), dy2.data_ptr(), N, P, Co, G, GN_EPS, d, st)
        # conv2: weight + data gradients
        conv3x3_wgrad(dy2, a1, gw2, N, H, W, Co, Co, Co)
        wk2t = conv_weight_dgrad(w2, dt)
        da1 = conv_s1(dy2, wk2t, N, H, W, Co, Co, dst=da2)                                # reuse da2 storage
        _call("pcm_gn_silu_bwd_reduce", da1.data_ptr(), 0, y1.data_ptr(), stats1.data_ptr(), g1.data_ptr(),
              b1.data_ptr(), gsum1.data_ptr(), gg1.data_ptr(), gb1.data_ptr(), N, P, Co, G, GN_EPS, d, st)
        dy1 = dy2                                                                          # reuse dy2 storage
        _call("pcm_gn_silu_bwd_apply", da1.data_ptr(), 0, y1.data_ptr(), stats1.data_ptr(), g1.data_ptr(),
              b1.data_ptr(), gsum1.data_ptr(), dy1.data_ptr(), N, P, Co, G, GN_EPS, d, st)
        conv3x3_wgrad(dy1, x, gw1, N, H, W, Co, Cip, Ci)
        dx = None
        if ctx.needs_input_grad[0]:
            wk1t = conv_weight_dgrad(w1, dt, Op=Cip)
            dx = conv_s1(dy1, wk1t, N, H, W, Co, Cip)
        return dx, rw1, rg1, rb1, rw2, rg2, rb2, rs1, rs2, rsp

    @staticmethod
    def _backward_fused(ctx, dout):
        x, y1, a1, y2, small, se, w1, g1, b1, w2, g2, b2, sw1, sw2, wsp, maps, ties, out = ctx.saved_tensors
        N, H, W, Cip, Ci, Co, Cr = ctx.dims
        G, dt = GN_GROUPS, x.dtype
        d, st = _DT[dt], _s()
        sdot = _take_sdot(dout)        # per-pixel sum_c dout*out, when the pooling backward that produced dout formed it
        dout = dout.contiguous()
        stats1, stats2, pool = small[: N * G * 2], small[N * G * 2: N * G * 4], small[N * G * 4:]
        hid = se[N * Co:]
        gw1, rw1 = _grad_buf(w1); gg1, rg1 = _grad_buf(g1); gb1, rb1 = _grad_buf(b1)
        gw2, rw2 = _grad_buf(w2); gg2, rg2 = _grad_buf(g2); gb2, rb2 = _grad_buf(b2)
        gs1, rs1 = _grad_buf(sw1); gs2, rs2 = _grad_buf(sw2); gsp, rsp = _grad_buf(wsp)
        dy2 = torch.empty_like(y2)
        if _GATE_WGRAD_SIDE:
            # the gate-weight gradient (98 sums per image, 11 % of the tail's instructions) leaves the critical path: the
            # tail writes dq, pcm_gate_wgrad accumulates dwsp for all images on the side stream
            dq = torch.empty(N * H * W, device=dout.device, dtype=torch.float32)
            _call("pcm_convblock_tail_bwd_sdot", dout.data_ptr(), y2.data_ptr(), out.data_ptr(), stats2.data_ptr(), g2.data_ptr(),
                  b2.data_ptr(), sw1.data_ptr(), sw2.data_ptr(), wsp.data_ptr(), pool.data_ptr(), se.data_ptr(),
                  hid.data_ptr(), maps.data_ptr(), ties.data_ptr(), dy2.data_ptr(), gg2.data_ptr(), gb2.data_ptr(),
                  gs1.data_ptr(), gs2.data_ptr(), gsp.data_ptr(), dq.data_ptr(), _p(sdot), N, H, W, Co, Cr, GN_EPS, d, st)
            with side_stream(dq, maps):
                _call("pcm_gate_wgrad", dq.data_ptr(), maps.data_ptr(), gsp.data_ptr(), N, H, W, _s())
        else:
            _call("pcm_convblock_tail_bwd_sdot", dout.data_ptr(), y2.data_ptr(), out.data_ptr(), stats2.data_ptr(), g2.data_ptr(),
                  b2.data_ptr(), sw1.data_ptr(), sw2.data_ptr(), wsp.data_ptr(), pool.data_ptr(), se.data_ptr(),
                  hid.data_ptr(), maps.data_ptr(), ties.data_ptr(), dy2.data_ptr(), gg2.data_ptr(), gb2.data_ptr(),
                  gs1.data_ptr(), gs2.data_ptr(), gsp.data_ptr(), 0, _p(sdot), N, H, W, Co, Cr, GN_EPS, d, st)
        with side_stream(dy2, a1):
            conv3x3_wgrad(dy2, a1, gw2, N, H, W, Co, Co, Co)
        gb = conv_group(dt, Co, Co, W)
        wk2t = conv_weight_dgrad(w2, dt, group=gb)
        da1 = conv_s1(dy2, wk2t, N, H, W, Co, Co, group=gb)
        dy1 = torch.empty_like(y1) if side_active() else dy2          # dy2 storage is reused unless a side kernel reads it
        _call("pcm_gn_silu_img_bwd", da1.data_ptr(), y1.data_ptr(), stats1.data_ptr(), g1.data_ptr(), b1.data_ptr(),
              dy1.data_ptr(), gg1.data_ptr(), gb1.data_ptr(), N, H, W, Co, GN_EPS, d, st)
        with side_stream(dy1, x):
            conv3x3_wgrad(dy1, x, gw1, N, H, W, Co, Cip, Ci)
        dx = None
        if ctx.needs_input_grad[0]:
            ga = conv_group(dt, Co, Cip, W)
            wk1t = conv_weight_dgrad(w1, dt, Op=Cip, group=ga)
            dx = conv_s1(dy1, wk1t, N, H, W, Co, Cip, group=ga)
        return dx, rw1, rg1, rb1, rw2, rg2, rb2, rs1, rs2, rsp


# ------------------------------------------------------------------------------------------------
# MaxPool2d(2) fused with the time-mean skip (src/unet_convlstm_attention.py:21,24,91-93)
# ------------------------------------------------------------------------------------------------
class PoolSkipFn(torch.autograd.Function):
    """s (B*T, H, W, C) -> pooled (B*T, H/2, W/2, C), skip (B, H, W, C) = mean over the T frames of
    each sample.  Backward is ONE kernel: max-pool routing (first max wins, like torch) + dskip/T."""

    @staticmethod
    def forward(ctx, s, T, up_channels=0):
        """s: t-major frames (image t*B + b).  up_channels > 0: the skip is written straight into the last C channels
        of a (B, H, W, up_channels + C) buffer — the concat buffer of the decoder stage that will consume it (UpCatFn
        recognises the view and fills the first up_channels channels in place: no copy of the skip)."""
        s = s.contiguous()
        N, H, W, C = s.shape
        B = N // T
        d = _DT[s.dtype]
        pooled = torch.empty((N, H // 2, W // 2, C), device=s.device, dtype=s.dtype)
        _call("pcm_maxpool2_fwd", s.data_ptr(), pooled.data_ptr(), N, H, W, C, d, _s())
        Cc = up_channels + C
        cat = torch.empty((B, H, W, Cc), device=s.device, dtype=s.dtype)
        skip = cat[..., up_channels:] if up_channels > 0 else cat
        with side_stream(s, cat):          # nothing needs the skip before the decoder (UpCatFn joins)
            _call("pcm_time_mean", s.data_ptr(), skip.data_ptr(), H * W * Cc, Cc, B, T, H * W, C, 1, d, _s())
        ctx.save_for_backward(s)
        ctx.T = T
        return pooled, skip

    @staticmethod
    def backward(ctx, dpooled, dskip):
        (s,) = ctx.saved_tensors
        N, H, W, C = s.shape
        ds = torch.empty_like(s)
        dp = dpooled.contiguous() if dpooled is not None else None
        ns = ps = 0
        if dskip is not None:
            # dskip may be a channel-slice view of the decoder's concat gradient: pass its strides
            assert dskip.stride(3) == 1 and dskip.stride(1) == W * dskip.stride(2)
            ns, ps = dskip.stride(0), dskip.stride(2)
        sdot = _offer_sdot(ds)
        _call("pcm_maxpool2_bwd_skip_dot", s.data_ptr(), _p(dp), _p(dskip), ns, ps, ds.data_ptr(), _p(sdot), N, H, W, C,
              ctx.T, 1, _DT[s.dtype], _s())
        return ds, None, None


class MaxPoolFn(torch.autograd.Function):
    """Plain nn.MaxPool2d(2) (src/unet.py:54)."""

    @staticmethod
    def forward(ctx, s):
        s = s.contiguous()
        N, H, W, C = s.shape
        pooled = torch.empty((N, H // 2, W // 2, C), device=s.device, dtype=s.dtype)
        _call("pcm_maxpool2_fwd", s.data_ptr(), pooled.data_ptr(), N, H, W, C, _DT[s.dtype], _s())
        ctx.save_for_backward(s)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        (s,) = ctx.saved_tensors
        N, H, W, C = s.shape
        ds = torch.empty_like(s)
        sdot = _offer_sdot(ds)
        _call("pcm_maxpool2_bwd_skip_dot", s.data_ptr(), dpooled.contiguous().data_ptr(), 0, 0, 0, ds.data_ptr(), _p(sdot),
              N, H, W, C, 1, 0, _DT[s.dtype], _s())
        return ds


# ------------------------------------------------------------------------------------------------
# Up: ConvTranspose2d(k2,s2,bias) + cat([up, skip])   (src/unet.py:63,67-68)
# ------------------------------------------------------------------------------------------------
class UpCatFn(torch.autograd.Function):
    """x (B,h,w,Ci), skip (B,2h,2w,Cs) -> cat (B,2h,2w,Co+Cs); the transposed conv writes straight
    into the first Co channels of the concat buffer."""

    @staticmethod
    def forward(ctx, x, skip, wt, bt):
        x = x.contiguous()
        B, h, w, Ci = x.shape
        Co, Cs = wt.shape[1], skip.shape[-1]
        assert wt.shape[0] == Ci and Co % 8 == 0 and Cs % 8 == 0
        H, W, Cc, dt = 2 * h, 2 * w, Co + Cs, x.dtype
        join_side()                        # the skip may have been produced on the side stream
        base = skip._base
        in_place = (base is not None and tuple(base.shape) == (B, H, W, Cc) and base.is_contiguous() and base.dtype == dt
                    and skip.data_ptr() == base.data_ptr() + Co * base.element_size()
                    and tuple(skip.stride()) == (H * W * Cc, W * Cc, Cc, 1))
        # PoolSkipFn(up_channels=Co) already put the skip into the last Cs channels of the concat buffer
        cat = base if in_place else torch.empty((B, H, W, Cc), device=x.device, dtype=dt)
        # transposed conv = one GEMM [pixels x Ci] x [Ci x 4*Co] whose epilogue pixel-shuffles into the concat buffer
        convT2x2_fwd(x, wt, bt, B, h, w, Ci, Co, cat, H * W * Cc, Cc)
        if not in_place:
            cat[..., Co:].copy_(skip)      # strided D2D copy (plumbing, no arithmetic)
        ctx.save_for_backward(x, wt, bt)
        ctx.dims = (B, h, w, Ci, Co, Cs)
        return cat

    @staticmethod
    def backward(ctx, dcat):
        x, wt, bt = ctx.saved_tensors
        B, h, w, Ci, Co, Cs = ctx.dims
        H, W, Cc, dt = 2 * h, 2 * w, Co + Cs, x.dtype
        dcat = dcat.contiguous()
        gwt, rwt = _grad_buf(wt)
        gbt, rbt = _grad_buf(bt)
        # data / weight gradients read the first Co channels of the concat gradient through stride-2 views
        dx = None
        if ctx.needs_input_grad[0]:
            dx = convT2x2_dgrad(dcat, wt, B, h, w, Ci, Co, H * W * Cc, Cc)
        with side_stream(x, dcat):
            convT2x2_wgrad(x, dcat, gwt, B, h, w, Ci, Co, H * W * Cc, Cc)
            channel_sum(dcat, gbt, B, H * W, Co, Co, ns=H * W * Cc, ps=Cc)
        dskip = dcat[..., Co:] if ctx.needs_input_grad[1] else None     # view; consumer reads it strided
        return dx, dskip, rwt, rbt


# ------------------------------------------------------------------------------------------------
# ConvLSTM (src/convlstm.py:11-35)
# ------------------------------------------------------------------------------------------------
def convlstm_seq_ok(H: int, W: int, Ch: int) -> bool:
    import os
    if os.environ.get("PCM_LSTM_PERSISTENT", "1") == "0":
        return False
    return bool(lib()._fn["pcm_convlstm_seq_supported"](H, W, Ch))


class ConvLSTMFn(torch.autograd.Function):
    """x: NHWC frames (T*B images); frame (t, b) is image t*st_t + b*st_b.  Returns h for all steps
    as (T, B, H, W, Ch), or only the last step when last_only (src/unet_convlstm_attention.py:88).

    W.cat(x,h) = Wx.x + Wh.h: the x half is computed for all T up front (it does not
    depend on the recurrence); each step then adds Wh.h_{t-1} and runs the fused cell update."""

    @staticmethod
    def forward(ctx, x, w, b, T, B, st_t, st_b, last_only):
        x = x.contiguous()
        _, H, W, Cip = x.shape
        Ch = w.shape[0] // 4
        Ci = w.shape[1] - Ch
        K = w.shape[-1]
        pad = K // 2
        assert Ci <= Cip and Cip % 8 == 0 and Ch % 8 == 0
        P, dt, dev = H * W, x.dtype, x.device
        d, st = _DT[dt], _s()
        img = P * Cip
        split = 256 if (dt == torch.bfloat16 and K == 3 and 4 * Ch > 256 and (4 * Ch) % 256 == 0) else 0
        wx = conv_weight_fwd(w, dt, 0, Ci, Ip=Cip, co_split=split)
        wh = conv_weight_fwd(w, dt, Ci, Ch, co_split=split)
        gates = torch.empty((T, B, P, 4 * Ch), device=dev, dtype=torch.float32)
        acts = torch.empty((T, B, P, 4 * Ch), device=dev, dtype=dt)
        c_all = torch.empty((T, B, P, Ch), device=dev, dtype=torch.float32)
        h_all = torch.empty((T, B, H, W, Ch), device=dev, dtype=dt)
        contiguous = (st_t == B and st_b == 1)          # t-major frames: every step is a block of B images
        if contiguous:
            conv_s1(x, wx, T * B, H, W, Cip, 4 * Ch, K, dst=gates, dst_f32=True, bias=b)      # Wx.x for ALL steps
        else:
            for t in range(T):
                conv_s1(x, wx, B, H, W, Cip, 4 * Ch, K, dst=gates[t], dst_f32=True, bias=b,
                        src_ns=st_b * img, src_off=t * st_t * img)
        fused = dt == torch.bfloat16 and K == 3 and Ch in (16, 32, 64)
        # all T steps in ONE launch (csrc/convlstm_seq.cu): clusters of four CTAs own two samples each, Wh resident in
        # shared memory, h exchanged through distributed shared memory.  PCM_LSTM_PERSISTENT=0 keeps one launch per step.
        persistent = (fused and contiguous and T > 1 and convlstm_seq_ok(H, W, Ch))
        ctx.persistent = persistent
        if persistent:
            _call("pcm_convlstm_seq_fwd_tc", gates.data_ptr(), wh.data_ptr(), h_all.data_ptr(), c_all.data_ptr(),
                  acts.data_ptr(), T, B, H, W, Ch, st)
        for t in range(T if not persistent else 0):
            if t > 0 and fused:
                # tcgen05 conv of h_{t-1} with the sigmoid/tanh cell fused in the epilogue (gates stay in TMEM)
                _call("pcm_convlstm_step_tc", h_all[t - 1].data_ptr(), wh.data_ptr(), gates[t].data_ptr(),
                      c_all[t - 1].data_ptr(), h_all[t].data_ptr(), c_all[t].data_ptr(), acts[t].data_ptr(),
                      B, H, W, Ch, st)
                continue
            if t > 0:
                conv_s1(h_all[t - 1], wh, B, H, W, Ch, 4 * Ch, K, dst=gates[t], dst_f32=True, accumulate=True)
            _call("pcm_lstm_cell_fwd", gates[t].data_ptr(), c_all[t - 1].data_ptr() if t > 0 else 0,
                  acts[t].data_ptr(), c_all[t].data_ptr(), h_all[t].data_ptr(), B * P, Ch, d, st)
        ctx.save_for_backward(x, w, b, acts, c_all, h_all)
        ctx.dims = (T, B, H, W, Cip, Ci, Ch, K, st_t, st_b, last_only)
        return h_all[T - 1] if last_only else h_all

    @staticmethod
    def backward(ctx, dh_ext):
        x, w, b, acts, c_all, h_all = ctx.saved_tensors
        T, B, H, W, Cip, Ci, Ch, K, st_t, st_b, last_only = ctx.dims
        pad = K // 2
        P, dt, dev = H * W, x.dtype, x.device
        d, st = _DT[dt], _s()
        img = P * Cip
        dh_ext = dh_ext.contiguous()
        gw, rw = _grad_buf(w)
        gb, rb = _grad_buf(b)
        dgates = torch.empty((T, B, H, W, 4 * Ch), device=dev, dtype=dt)
        dc = [torch.empty((B, P, Ch), device=dev, dtype=torch.float32) for _ in range(2)]
        wht = conv_weight_dgrad(w, dt, Ci, Ch)
        dh_next = None
        if ctx.persistent:
            # back-propagation through time in ONE launch: cell backward in registers, Wh^T.dgates split over K across
            # the four CTAs of a cluster, partial sums exchanged through distributed shared memory
            _call("pcm_convlstm_seq_bwd_tc", dh_ext.data_ptr(), 0 if last_only else 1, acts.data_ptr(), c_all.data_ptr(),
                  wht.data_ptr(), dgates.data_ptr(), T, B, H, W, Ch, st)
        for t in range(T - 1, -1, -1) if not ctx.persistent else ():
            if last_only:
                ext = dh_ext if t == T - 1 else None
            else:
                ext = dh_ext[t]
            _call("pcm_lstm_cell_bwd", _p(ext), _p(dh_next), dc[(t + 1) & 1].data_ptr() if t < T - 1 else 0,
                  acts[t].data_ptr(), c_all[t - 1].data_ptr() if t > 0 else 0, c_all[t].data_ptr(),
                  dgates[t].data_ptr(), dc[t & 1].data_ptr(), B * P, Ch, d, st)
            if t > 0:
                dh_next = conv_s1(dgates[t], wht, B, H, W, 4 * Ch, Ch, K, dst=dh_next)
        KK = K * K
        Ct = Ci + Ch
        # dW[:, :Ci] — x frames may be time-strided, one launch per step; dW[:, Ci:] — one launch over t>=1
        contiguous = (st_t == B and st_b == 1)
        if K == 3:
          with side_stream(dgates, x, h_all):       # the weight / bias gradients overlap the dx convolution below
            if contiguous:
                conv3x3_wgrad(dgates, x, gw, T * B, H, W, 4 * Ch, Cip, Ci, Ci_tot=Ct)
            else:
                for t in range(T):
                    conv3x3_wgrad(dgates[t], x, gw, B, H, W, 4 * Ch, Cip, Ci, Ci_tot=Ct, x_ns=st_b * img,
                                  x_off=t * st_t * img)
            if T > 1:
                conv3x3_wgrad(dgates[1:], h_all[:-1], gw, (T - 1) * B, H, W, 4 * Ch, Ch, Ch, Ci_tot=Ct, dw_off=Ci * KK)
            channel_sum(dgates, gb, T * B, P, 4 * Ch, 4 * Ch)
        else:
            for t in range(T):
                conv_wgrad(dgates[t], x, gw, Ct * KK, KK, 1, B, H, W, 4 * Ch, 4 * Ch, H, W, Cip, Ci, K, K, 1, pad,
                           b_ns=st_b * img, b_off=t * st_t * img)
            if T > 1:
                conv_wgrad(dgates[1:], h_all[:-1], gw, Ct * KK, KK, 1, (T - 1) * B, H, W, 4 * Ch, 4 * Ch, H, W, Ch, Ch,
                           K, K, 1, pad, dw_off=Ci * KK)
            channel_sum(dgates, gb, T * B, P, 4 * Ch, 4 * Ch)
        dx = None
        if ctx.needs_input_grad[0]:
            wxt = conv_weight_dgrad(w, dt, 0, Ci, Op=Cip)
            dx = torch.empty_like(x)
            if contiguous:
                conv_s1(dgates, wxt, T * B, H, W, 4 * Ch, Cip, K, dst=dx)
            else:
                for t in range(T):
                    conv_s1(dgates[t], wxt, B, H, W, 4 * Ch, Cip, K, dst=dx, dst_ns=st_b * img, dst_off=t * st_t * img)
        return dx, rw, rb, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# head (1x1 conv + bias -> NCHW fp32) and MSE loss
# ------------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    """nn.Conv2d(base, out_ch, 1) (src/unet_convlstm_attention.py:56,104): NHWC in, NCHW fp32 out."""

    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        N, H, W, C = x.shape
        K = w.shape[0]
        out = torch.empty((N, K, H, W), device=x.device, dtype=torch.float32)
        _call("pcm_head_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), N, H * W, C, K, _DT[x.dtype], _s())
        ctx.save_for_backward(x, w, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, b = ctx.saved_tensors
        N, H, W, C = x.shape
        K = w.shape[0]
        dout = dout.contiguous().float()
        gw, rw = _grad_buf(w)
        gb, rb = _grad_buf(b)
        dx = torch.empty_like(x)
        _call("pcm_head_bwd", dout.data_ptr(), x.data_ptr(), w.data_ptr(), dx.data_ptr(), gw.data_ptr(), gb.data_ptr(),
              N, H * W, C, K, _DT[x.dtype], _s())
        return dx, rw, rb


class HeadMSEFn(torch.autograd.Function):
    """loss = nn.MSELoss()(head(x), target) with the 1x1 head (src/unet_convlstm_attention.py:104, main_final.py:559) in
    two launches instead of four: the prediction never goes to HBM in training (the step needs it only inside the loss)."""

    @staticmethod
    def forward(ctx, x, w, b, target):
        x = x.contiguous()
        target = target.contiguous().float()
        N, H, W, C = x.shape
        K = w.shape[0]
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        _call("pcm_head_mse_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), target.data_ptr(), 0, loss.data_ptr(), N, H * W, C, K,
              _DT[x.dtype], _s())
        ctx.save_for_backward(x, w, b, target)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, w, b, target = ctx.saved_tensors
        N, H, W, C = x.shape
        K = w.shape[0]
        g = g.contiguous().float()
        gw, rw = _grad_buf(w)
        gb, rb = _grad_buf(b)
        dx = torch.empty_like(x)
        _call("pcm_head_mse_bwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), target.data_ptr(), g.data_ptr(), dx.data_ptr(),
              gw.data_ptr(), gb.data_ptr(), N, H * W, C, K, _DT[x.dtype], _s())
        return dx, rw, rb, None


def head_mse_ok(C: int, K: int) -> bool:
    import os
    return os.environ.get("PCM_HEAD_MSE", "1") != "0" and bool(lib()._fn["pcm_head_mse_supported"](C, K))


def head_or_loss(x, w, b, target=None):
    """The model's final 1x1 convolution; with `target` the training loss nn.MSELoss()(head(x), target) instead — fused
    (HeadMSEFn) when the head shape allows, else head followed by the loss."""
    if target is None:
        return HeadFn.apply(x, w, b)
    if head_mse_ok(x.shape[-1], w.shape[0]):
        return HeadMSEFn.apply(x, w, b, target)
    return mse_loss(HeadFn.apply(x, w, b), target)


class MSELossFn(torch.autograd.Function):
    """nn.MSELoss() (main_final.py:544,559)."""

    @staticmethod
    def forward(ctx, pred, target):
        _require_cuda(pred, "prediction")
        pred = pred.contiguous().float()
        target = target.contiguous().float()
        loss = torch.zeros(1, device=pred.device, dtype=torch.float32)
        _call("pcm_mse_fwd", pred.data_ptr(), target.data_ptr(), loss.data_ptr(), pred.numel(), _s())
        ctx.save_for_backward(pred, target)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        g = g.contiguous().float()
        dp = torch.empty_like(pred)
        _call("pcm_mse_bwd", pred.data_ptr(), target.data_ptr(), g.data_ptr(), dp.data_ptr(), pred.numel(), _s())
        return dp, None


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    return MSELossFn.apply(pred, target)


# ------------------------------------------------------------------------------------------------
# stand-alone attention helpers and single cell step (reference modules usable on their own)
# ------------------------------------------------------------------------------------------------
class SEFn(torch.autograd.Function):
    """SEBlock.forward (src/unet.py:16-17): x * sigmoid(W2 relu(W1 avgpool(x))) on NHWC."""

    @staticmethod
    def forward(ctx, a, w1, w2):
        a = a.contiguous()
        N, H, W, C = a.shape
        if w1.shape[1] != C:
            raise RuntimeError("pcm_b200 SEBlock needs a channel count that is a multiple of 8")
        Cr, P, d, st, dev = w1.shape[0], H * W, _DT[a.dtype], _s(), a.device
        pool = torch.zeros(N * C, device=dev, dtype=torch.float32)
        channel_sum(a, pool, N, P, C, C, per_image=True)
        se = torch.empty(N * C + N * Cr, device=dev, dtype=torch.float32)
        hid = se[N * C:]
        cmap = torch.empty(N * P * 2, device=dev, dtype=torch.float32)
        _call("pcm_se_chanstat_fwd", a.data_ptr(), pool.data_ptr(), w1.data_ptr(), w2.data_ptr(), se.data_ptr(),
              hid.data_ptr(), cmap.data_ptr(), N, P, C, Cr, d, st)
        out = torch.empty_like(a)
        _call("pcm_scale_channels", a.data_ptr(), se.data_ptr(), 0, out.data_ptr(), N, P, C, d, st)
        ctx.save_for_backward(a, pool, se, w1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, pool, se, w1, w2 = ctx.saved_tensors
        N, H, W, C = a.shape
        Cr, P, d, st, dev = w1.shape[0], H * W, _DT[a.dtype], _s(), a.device
        dout = dout.contiguous()
        hid = se[N * C:]
        g1, r1 = _grad_buf(w1)
        g2, r2 = _grad_buf(w2)
        # reuse the gate backward with gate = 1, dq = 0: da = dout*se, dse = sum_p dout*a
        ones = torch.ones(N * P, device=dev, dtype=torch.float32)
        zer = torch.zeros(N * P * 3 + 98 + N * C, device=dev, dtype=torch.float32)
        dq, cmap, wsp0, dse = zer[: N * P], zer[N * P: N * P * 3], zer[N * P * 3: N * P * 3 + 98], zer[N * P * 3 + 98:]
        da = torch.empty_like(a)
        _call("pcm_spatial_gate_bwd_da", dout.data_ptr(), a.data_ptr(), se.data_ptr(), ones.data_ptr(), cmap.data_ptr(),
              dq.data_ptr(), wsp0.data_ptr(), da.data_ptr(), dse.data_ptr(), N, H, W, C, d, st)
        dpool = torch.empty(N * C, device=dev, dtype=torch.float32)
        _call("pcm_se_bwd", dse.data_ptr(), se.data_ptr(), hid.data_ptr(), pool.data_ptr(), w1.data_ptr(), w2.data_ptr(),
              dpool.data_ptr(), g1.data_ptr(), g2.data_ptr(), N, P, C, Cr, st)
        _call("pcm_scale_channels", da.data_ptr(), 0, dpool.data_ptr(), da.data_ptr(), N, P, C, d, st)
        return da, r1, r2


class SpatialGateFn(torch.autograd.Function):
    """SpatialGate.forward (src/unet.py:25-29) on NHWC: x * sigmoid(conv7x7([mean_C x, amax_C x]))."""

    @staticmethod
    def forward(ctx, a, wsp):
        a = a.contiguous()
        N, H, W, C = a.shape
        P, d, st, dev = H * W, _DT[a.dtype], _s(), a.device
        se = torch.ones(N * C + N, device=dev, dtype=torch.float32)      # se = 1 (no excitation); hid dummy
        maps = torch.empty(N * P * 3, device=dev, dtype=torch.float32)
        cmap, gate = maps[: N * P * 2], maps[N * P * 2:]
        _call("pcm_se_chanstat_fwd", a.data_ptr(), 0, 0, 0, se.data_ptr(), se[N * C:].data_ptr(), cmap.data_ptr(),
              N, P, C, 1, d, st)
        out = torch.empty_like(a)
        _call("pcm_spatial_gate_fwd", a.data_ptr(), se.data_ptr(), cmap.data_ptr(), wsp.data_ptr(), gate.data_ptr(),
              out.data_ptr(), N, H, W, C, d, st)
        ctx.save_for_backward(a, se, maps, wsp)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, se, maps, wsp = ctx.saved_tensors
        N, H, W, C = a.shape
        P, d, st, dev = H * W, _DT[a.dtype], _s(), a.device
        dout = dout.contiguous()
        cmap, gate = maps[: N * P * 2], maps[N * P * 2:]
        gsp, rsp = _grad_buf(wsp)
        dq = torch.empty(N * P, device=dev, dtype=torch.float32)
        dse = torch.zeros(N * C, device=dev, dtype=torch.float32)
        _call("pcm_spatial_gate_bwd_dq", dout.data_ptr(), a.data_ptr(), se.data_ptr(), gate.data_ptr(), dq.data_ptr(),
              N, P, C, d, st)
        _call("pcm_spatial_gate_bwd_dw", dq.data_ptr(), cmap.data_ptr(), gsp.data_ptr(), N, H, W, st)
        da = torch.empty_like(a)
        _call("pcm_spatial_gate_bwd_da", dout.data_ptr(), a.data_ptr(), se.data_ptr(), gate.data_ptr(), cmap.data_ptr(),
              dq.data_ptr(), wsp.data_ptr(), da.data_ptr(), dse.data_ptr(), N, H, W, C, d, st)
        return da, rsp


class CellStepFn(torch.autograd.Function):
    """ConvLSTMCell.forward (src/convlstm.py:11-19) for one explicit step: xh = NHWC cat([x, h]),
    c = NHWC fp32 cell state -> (h_next NHWC, c_next NHWC fp32)."""

    @staticmethod
    def forward(ctx, xh, c, w, b):
        xh, c = xh.contiguous(), c.contiguous()
        B, H, W, Cp = xh.shape
        Ch, Ct, K = w.shape[0] // 4, w.shape[1], w.shape[-1]
        assert Ct <= Cp and Ch % 8 == 0
        P, dt, dev = H * W, xh.dtype, xh.device
        wk = conv_weight_fwd(w, dt, Ip=Cp)
        gates = conv_s1(xh, wk, B, H, W, Cp, 4 * Ch, K, dst_f32=True, bias=b)
        acts = torch.empty((B, P, 4 * Ch), device=dev, dtype=dt)
        c2 = torch.empty((B, H, W, Ch), device=dev, dtype=torch.float32)
        h2 = torch.empty((B, H, W, Ch), device=dev, dtype=dt)
        _call("pcm_lstm_cell_fwd", gates.data_ptr(), c.data_ptr(), acts.data_ptr(), c2.data_ptr(), h2.data_ptr(),
              B * P, Ch, _DT[dt], _s())
        ctx.save_for_backward(xh, c, w, b, acts, c2)
        return h2, c2

    @staticmethod
    def backward(ctx, dh, dc):
        xh, c, w, b, acts, c2 = ctx.saved_tensors
        B, H, W, Cp = xh.shape
        Ch, Ct, K = w.shape[0] // 4, w.shape[1], w.shape[-1]
        P, dt, dev = H * W, xh.dtype, xh.device
        gw, rw = _grad_buf(w)
        gb, rb = _grad_buf(b)
        dgates = torch.empty((B, H, W, 4 * Ch), device=dev, dtype=dt)
        dcp = torch.empty((B, H, W, Ch), device=dev, dtype=torch.float32)
        _call("pcm_lstm_cell_bwd", _p(dh.contiguous() if dh is not None else None), 0,
              _p(dc.contiguous().float() if dc is not None else None), acts.data_ptr(), c.data_ptr(), c2.data_ptr(),
              dgates.data_ptr(), dcp.data_ptr(), B * P, Ch, _DT[dt], _s())
        conv_wgrad(dgates, xh, gw, Ct * K * K, K * K, 1, B, H, W, 4 * Ch, 4 * Ch, H, W, Cp, Ct, K, K, 1, K // 2)
        channel_sum(dgates, gb, B, P, 4 * Ch, 4 * Ch)
        wkt = conv_weight_dgrad(w, dt, Op=Cp)
        dxh = conv_s1(dgates, wkt, B, H, W, 4 * Ch, Cp, K)
        return dxh, dcp, rw, rb
