"""Micro-benchmarks of individual pcm_b200 kernels at the config-3 shapes (B=64, T=6, 48x72, base 16).
Timing: CUDA events around `iters` back-to-back launches after warm-up; a 256 MB buffer is rewritten
between launches (--flush) so operands do not stay L2 resident.

    python tools/microbench.py conv [--flush] [--only NAME]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcm_b200  # noqa: E402,F401
from pcm_b200 import ops  # noqa: E402

CONV_SHAPES = [  # name, N, H, W, Cin, Cout
    ("enc1.body.0", 384, 48, 72, 16, 16), ("enc1.body.3", 384, 48, 72, 16, 16),
    ("enc2.body.0", 384, 24, 36, 16, 32), ("enc2.body.3", 384, 24, 36, 32, 32),
    ("enc3.body.0", 384, 12, 18, 32, 64), ("enc3.body.3", 384, 12, 18, 64, 64),
    ("enc4.body.0", 384, 6, 9, 64, 128), ("enc4.body.3", 384, 6, 9, 128, 128),
    ("lstm.Wx(all T)", 384, 6, 9, 128, 256), ("lstm.Wh(step)", 64, 6, 9, 64, 256),
    ("up3.body.0", 64, 12, 18, 128, 64), ("up2.body.0", 64, 24, 36, 64, 32), ("up1.body.0", 64, 48, 72, 32, 16),
]


def timeit(fn, iters, flush):
    """flush: one launch per measurement with a 256 MB L2 flush in between (includes ~host launch gap);
    otherwise: `iters` launches captured in ONE CUDA graph and replayed (pure device time, warm L2)."""
    buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda") if flush else None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if not flush:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    tot = 0.0
    for _ in range(iters):
        if flush:
            buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def bench_conv(args):
    for name, N, H, W, Ci, Co in CONV_SHAPES:
        if args.only and args.only not in name:
            continue
        x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
        w = torch.randn(Co, Ci, 3, 3, device="cuda") / (3 * Ci ** 0.5)
        grp = ops.conv_group(torch.bfloat16, Ci, Co, W)          # PCM_CONV_GROUP / PCM_CONV_GROUP_MAXC select the form
        wk = ops.conv_weight_fwd(w, torch.bfloat16, group=grp)
        y = torch.empty(N, H, W, Co, device="cuda", dtype=torch.bfloat16)
        name = f"{name} g{grp}"
        ms = timeit(lambda: ops.conv_s1(x, wk, N, H, W, Ci, Co, dst=y, group=grp), args.iters, args.flush)
        flops = 2.0 * N * H * W * Ci * Co * 9
        byts = 2.0 * N * H * W * (Ci + Co)
        print(f"conv3x3 {name:16s} N={N:4d} {H:2d}x{W:2d} {Ci:3d}->{Co:3d}: {ms * 1e3:8.1f} us  "
              f"{flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s (algorithmic)")


def bench_wgrad(args):
    for name, N, H, W, Ci, Co in CONV_SHAPES:
        if args.only and args.only not in name:
            continue
        x = torch.randn(N, H, W, Ci, device="cuda").bfloat16()
        dy = torch.randn(N, H, W, Co, device="cuda").bfloat16()
        dw = torch.zeros(Co, Ci, 3, 3, device="cuda")
        ms = timeit(lambda: ops.conv3x3_wgrad(dy, x, dw, N, H, W, Co, Ci, Ci), args.iters, args.flush)
        flops = 2.0 * N * H * W * Ci * Co * 9
        byts = 2.0 * N * H * W * (Ci + Co)
        print(f"wgrad   {name:16s} N={N:4d} {H:2d}x{W:2d} {Ci:3d}->{Co:3d}: {ms * 1e3:8.1f} us  "
              f"{flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s (algorithmic)")


def bench_pointwise(args):
    """ConvBlock tail kernels at the enc1 level (N=384, 48x72, C=16) and enc3 level (12x18, C=64)."""
    from pcm_b200.ops import _call, _s
    for (N, H, W, C) in [(384, 48, 72, 16), (384, 24, 36, 32), (384, 12, 18, 64), (384, 6, 9, 128)]:
        P, G, d = H * W, 8, 1
        Cr = C // 8
        bf = lambda *sh: torch.randn(*sh, device="cuda").bfloat16()
        x, da, y, dout = bf(N, P, C), bf(N, P, C), bf(N, P, C), bf(N, P, C)
        f32 = lambda *sh: torch.randn(*sh, device="cuda")
        stats = torch.zeros(N * G * 2, device="cuda")
        gamma, beta = f32(C), f32(C)
        pool, se, hid = torch.zeros(N * C, device="cuda"), torch.rand(N * C, device="cuda"), torch.rand(N * Cr, device="cuda")
        w1, w2, wsp = f32(Cr * C), f32(C * Cr), f32(98)
        cmap, gate, dq = f32(N * P * 2), torch.rand(N * P, device="cuda"), f32(N * P)
        gsum, dg, db, dse, dwsp = (torch.zeros(n, device="cuda") for n in (N * G * 2, C, C, N * C, 98))
        _call("pcm_gn_stats", x.data_ptr(), stats.data_ptr(), N, P, C, G, d, _s())
        cases = [
            ("gn_stats", 1, lambda: _call("pcm_gn_stats", x.data_ptr(), stats.data_ptr(), N, P, C, G, d, _s())),
            ("gn_silu_fwd(+pool)", 2, lambda: _call("pcm_gn_silu_fwd", x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), pool.data_ptr(), N, P, C, G, 1e-5, d, _s())),
            ("se_chanstat_fwd", 1, lambda: _call("pcm_se_chanstat_fwd", y.data_ptr(), pool.data_ptr(), w1.data_ptr(), w2.data_ptr(), se.data_ptr(), hid.data_ptr(), cmap.data_ptr(), N, P, C, Cr, d, _s())),
            ("spatial_gate_fwd", 2, lambda: _call("pcm_spatial_gate_fwd", y.data_ptr(), se.data_ptr(), cmap.data_ptr(), wsp.data_ptr(), gate.data_ptr(), x.data_ptr(), N, H, W, C, d, _s())),
            ("spatial_gate_bwd_dq", 2, lambda: _call("pcm_spatial_gate_bwd_dq", dout.data_ptr(), y.data_ptr(), se.data_ptr(), gate.data_ptr(), dq.data_ptr(), N, P, C, d, _s())),
            ("spatial_gate_bwd_dw", 0, lambda: _call("pcm_spatial_gate_bwd_dw", dq.data_ptr(), cmap.data_ptr(), dwsp.data_ptr(), N, H, W, _s())),
            ("spatial_gate_bwd_da", 3, lambda: _call("pcm_spatial_gate_bwd_da", dout.data_ptr(), y.data_ptr(), se.data_ptr(), gate.data_ptr(), cmap.data_ptr(), dq.data_ptr(), wsp.data_ptr(), da.data_ptr(), dse.data_ptr(), N, H, W, C, d, _s())),
            ("gn_silu_bwd_reduce", 2, lambda: _call("pcm_gn_silu_bwd_reduce", da.data_ptr(), pool.data_ptr(), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), gsum.data_ptr(), dg.data_ptr(), db.data_ptr(), N, P, C, G, 1e-5, d, _s())),
            ("gn_silu_bwd_apply", 3, lambda: _call("pcm_gn_silu_bwd_apply", da.data_ptr(), pool.data_ptr(), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), gsum.data_ptr(), y.data_ptr(), N, P, C, G, 1e-5, d, _s())),
        ]
        for name, ntens, fn in cases:
            if args.only and args.only not in name:
                continue
            ms = timeit(fn, args.iters, args.flush)
            byts = ntens * N * P * C * 2.0
            print(f"{name:22s} N={N} {H}x{W} C={C:3d}: {ms * 1e3:8.1f} us   {byts / ms / 1e6:8.1f} GB/s (algorithmic, {ntens} tensor passes)")


def bench_tails(args):
    """Per-image fused ConvBlock tails (csrc/convblock_fused.cu) at the four encoder levels and the decoder's top level."""
    from pcm_b200.ops import _call, _s
    for (N, H, W, C) in [(384, 48, 72, 16), (384, 24, 36, 32), (384, 12, 18, 64), (384, 6, 9, 128), (64, 48, 72, 16)]:
        P, d, Cr = H * W, 1, C // 8
        bf = lambda *sh: torch.randn(*sh, device="cuda").bfloat16()
        f32 = lambda *sh: torch.randn(*sh, device="cuda")
        x, y, dout, dx = bf(N, P, C), bf(N, P, C), bf(N, P, C), bf(N, P, C)
        gamma, beta = f32(C), f32(C)
        w1, w2, wsp = f32(Cr * C) / C ** 0.5, f32(C * Cr), f32(98) / 7
        stats, pool, se, hid = torch.zeros(N * 16, device="cuda"), torch.zeros(N * C, device="cuda"), torch.zeros(N * C, device="cuda"), torch.zeros(N * Cr, device="cuda")
        dg, db, dw1, dw2, dwsp = (torch.zeros(n, device="cuda") for n in (C, C, Cr * C, C * Cr, 98))
        maps, ties = torch.zeros(N * P * 3, device="cuda"), torch.zeros(N * P, device="cuda", dtype=torch.uint8)
        cases = [
            ("gn_silu_img_fwd", 2, lambda: _call("pcm_gn_silu_img_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), stats.data_ptr(), y.data_ptr(), N, H, W, C, 1e-5, d, _s())),
            ("convblock_tail_fwd", 2, lambda: _call("pcm_convblock_tail_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), w2.data_ptr(), wsp.data_ptr(), stats.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(), maps.data_ptr(), ties.data_ptr(), y.data_ptr(), N, H, W, C, Cr, 1e-5, d, _s())),
            ("gn_silu_img_bwd", 3, lambda: _call("pcm_gn_silu_img_bwd", dout.data_ptr(), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), N, H, W, C, 1e-5, d, _s())),
            ("convblock_tail_bwd", 4, lambda: _call("pcm_convblock_tail_bwd", dout.data_ptr(), x.data_ptr(), y.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), w2.data_ptr(), wsp.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(), maps.data_ptr(), ties.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dw1.data_ptr(), dw2.data_ptr(), dwsp.data_ptr(), N, H, W, C, Cr, 1e-5, d, _s())),
        ]
        sdot = torch.randn(N * P, device="cuda")
        cases.append(("convblock_tail_bwd_sdot", 3, lambda: _call("pcm_convblock_tail_bwd_sdot", dout.data_ptr(), x.data_ptr(), y.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), w2.data_ptr(), wsp.data_ptr(), pool.data_ptr(), se.data_ptr(), hid.data_ptr(), maps.data_ptr(), ties.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dw1.data_ptr(), dw2.data_ptr(), dwsp.data_ptr(), 0, sdot.data_ptr(), N, H, W, C, Cr, 1e-5, d, _s())))
        cases[1][2]()                      # populate stats / pool / se / hid for the backward kernels
        for name, ntens, fn in cases:
            if args.only and args.only not in name:
                continue
            ms = timeit(fn, args.iters, args.flush)
            byts = ntens * N * P * C * 2.0
            print(f"{name:22s} N={N} {H}x{W} C={C:3d}: {ms * 1e3:8.1f} us   {byts / ms / 1e6:8.1f} GB/s (algorithmic, {ntens} tensor passes)   "
                  f"{ms * 1e6 / (N * P * C):7.3f} ns/element")


def bench_mha(args):
    """Attention core of CNNTransformer (B=64, L=216, 4 heads of 32): tcgen05 vs SIMT kernels, forward and backward."""
    from pcm_b200.ops import _call, _s
    B, L, nh, D = 64, 216, 4, 32
    E = nh * D
    qkv = torch.randn(B, L, 3 * E, device="cuda").bfloat16()
    dout = (torch.randn(B, L, E, device="cuda") / 8).bfloat16()
    out = torch.empty(B, L, E, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * nh * L, device="cuda")
    dqkv = torch.empty_like(qkv)
    sc = 1.0 / D ** 0.5
    for pd in (0.0, 0.1):
        cases = [
            ("mha_fwd_tc", 1.0, lambda: _call("pcm_mha_fwd_tc", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, nh, sc, pd, 7, _s())),
            ("mha_fwd (SIMT)", 1.0, lambda: _call("pcm_mha_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, nh, D, sc, pd, 7, 1, _s())),
            ("mha_bwd_tc", 2.5, lambda: _call("pcm_mha_bwd_tc", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, L, nh, sc, pd, 7, _s())),
            ("mha_bwd (SIMT)", 2.5, lambda: _call("pcm_mha_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, L, nh, D, sc, pd, 7, 1, _s())),
        ]
        for name, k, fn in cases:
            if args.only and args.only not in name:
                continue
            ms = timeit(fn, args.iters, args.flush)
            flops = 4.0 * B * nh * L * L * D * k
            print(f"{name:18s} dropout {pd}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.2f} TFLOP/s")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["conv", "wgrad", "pointwise", "tails", "mha"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    {"conv": bench_conv, "wgrad": bench_wgrad, "pointwise": bench_pointwise, "tails": bench_tails, "mha": bench_mha}[a.what](a)
