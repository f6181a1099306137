"""Algorithmic cost of one C-ABI call: (FLOPs, HBM bytes) from the call's own arguments.

Used by bench.py for the per-kernel roofline (`achieved` = algorithmic work / measured duration) and by
tools/microbench.py.  "Algorithmic" = what the operation must move/compute at minimum: every input element
read once, every output written once, MACs x 2 — not what the implementation happens to do.
Argument names are those of include/pcm_b200.h (the binding parses them from the header)."""
from __future__ import annotations

from ._lib import PROTOS


def _named(name, args):
    return {an: a for (_, an), a in zip(PROTOS[name][1], args)}


def algo_cost(name: str, args):
    """-> (flops, bytes, bound) with bound in {"tensor", "hbm"}; (0, 0, "hbm") for unknown calls."""
    a = _named(name, args)
    es = lambda: 2 if a.get("dtype", 1) == 1 else 4          # activation element size
    if name in ("pcm_conv3x3_tc", "pcm_conv1x1_tc"):
        px = a["N"] * a["H"] * a["W"]
        taps = 9 if name == "pcm_conv3x3_tc" else 1
        fl = 2.0 * px * a["Cin"] * a["Cout"] * taps
        osz = 4 if a["dst_f32"] else 2
        by = px * (a["Cin"] * 2 + a["Cout"] * osz * (2 if a["accumulate"] else 1)) + taps * a["Cin"] * a["Cout"] * 2
        return fl, by, "tensor"
    if name == "pcm_conv3x3_tc_grouped":               # algorithmic cost of the layer, not of the zero-padded group GEMM
        px = a["N"] * a["H"] * a["W"]
        fl = 2.0 * px * a["Cin"] * a["Cout"] * 9
        by = px * (a["Cin"] * 2 + a["Cout"] * (4 if a["dst_f32"] else 2)) + 9 * a["Cin"] * a["Cout"] * 2
        return fl, by, "tensor"
    if name == "pcm_convlstm_step_tc":
        px = a["B"] * a["H"] * a["W"]
        Ch = a["Ch"]
        fl = 2.0 * px * Ch * 4 * Ch * 9
        # h_prev bf16 + gx fp32 [4Ch] + c_prev fp32 in; h bf16 + c fp32 + acts bf16 [4Ch] out
        by = px * (Ch * 2 + 4 * Ch * 4 + Ch * 4 + Ch * 2 + Ch * 4 + 4 * Ch * 2) + 9 * Ch * 4 * Ch * 2
        return fl, by, "tensor"
    if name == "pcm_convlstm_seq_fwd_tc":
        # T-1 recurrent gate convolutions (h_0 = 0) ; per step: gx fp32 [4Ch] in, acts bf16 [4Ch] + c fp32 + h bf16 out
        px = a["B"] * a["H"] * a["W"]
        Ch, T = a["Ch"], a["T"]
        fl = 2.0 * px * Ch * 4 * Ch * 9 * (T - 1)
        by = T * px * (4 * Ch * 4 + 4 * Ch * 2 + Ch * 4 + Ch * 2) + 9 * Ch * 4 * Ch * 2
        return fl, by, "tensor"
    if name == "pcm_convlstm_seq_bwd_tc":
        # T-1 data-gradient convolutions ; per step: acts bf16 [4Ch] + c (twice) fp32 in, dgates bf16 [4Ch] out
        px = a["B"] * a["H"] * a["W"]
        Ch, T = a["Ch"], a["T"]
        fl = 2.0 * px * Ch * 4 * Ch * 9 * (T - 1)
        by = T * px * (4 * Ch * 2 + 2 * Ch * 4 + 4 * Ch * 2) + px * Ch * 2 + 9 * Ch * 4 * Ch * 2
        return fl, by, "tensor"
    if name in ("pcm_wgrad3x3_tc", "pcm_wgrad1x1_tc"):
        px = a["N"] * a["H"] * a["W"]
        taps = 9 if name == "pcm_wgrad3x3_tc" else 1
        fl = 2.0 * px * a["Co"] * a["Ci"] * taps
        by = px * (a["Co"] + a["Ci"]) * 2 + taps * a["Co_real"] * a["Ci_real"] * 4
        return fl, by, "tensor"
    if name == "pcm_wgrad3x3_tc_grouped":
        px = a["N"] * a["H"] * a["W"]
        return 2.0 * px * a["Co"] * a["Ci"] * 9, px * (a["Co"] + a["Ci"]) * 2 + 9 * a["Co"] * a["Ci_real"] * 4, "tensor"
    if name == "pcm_convT2x2_tc":          # H, W = input grid; output has 4x the pixels
        px = a["N"] * a["H"] * a["W"]
        return 2.0 * px * a["Cin"] * 4 * a["Cout"], px * (a["Cin"] + 4 * a["Cout"]) * 2 + 4 * a["Cin"] * a["Cout"] * 2, "tensor"
    if name == "pcm_convT2x2_dgrad_tc":
        px = a["N"] * a["H"] * a["W"]
        return 2.0 * px * a["Cin"] * 4 * a["Cout"], px * (a["Cin"] + 4 * a["Cout"]) * 2 + 4 * a["Cin"] * a["Cout"] * 2, "tensor"
    if name == "pcm_convT2x2_wgrad_tc":
        px = a["N"] * a["H"] * a["W"]
        return 2.0 * px * a["Ca"] * 4 * a["Cb"], px * (a["Ca"] + 4 * a["Cb"]) * 2 + 4 * a["Ca_real"] * a["Cb_real"] * 4, "tensor"
    # per-image fused ConvBlock tails: every tensor once (x in, y out; backward: dout + x in, dx out); the second
    # tail also writes / reads the saved gate maps (3 fp32 + 1 byte per pixel) and its backward reads `out`
    if name == "pcm_gn_silu_img_fwd":
        return 0.0, 2 * a["N"] * a["H"] * a["W"] * a["C"] * es(), "hbm"
    if name == "pcm_convblock_tail_fwd":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, 2 * px * a["C"] * es() + (13 * px if a["maps"] else 0), "hbm"
    if name == "pcm_gn_silu_img_bwd":
        return 0.0, 3 * a["N"] * a["H"] * a["W"] * a["C"] * es(), "hbm"
    if name == "pcm_gate_wgrad":
        return 2.0 * 98 * a["N"] * a["H"] * a["W"], a["N"] * a["H"] * a["W"] * 12, "hbm"
    if name == "pcm_convblock_tail_bwd_dq":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, 4 * px * a["C"] * es() + 17 * px, "hbm"
    if name == "pcm_convblock_tail_bwd":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, 4 * px * a["C"] * es() + 13 * px, "hbm"
    if name == "pcm_convblock_tail_bwd_sdot":
        # with sdot (4 bytes per pixel, formed by the pooling backward) neither `out` nor a second pass over dout is part
        # of the algorithm: dout + x in, dx out; without it this is pcm_convblock_tail_bwd(_dq)
        px = a["N"] * a["H"] * a["W"]
        tensors = 3 if a["sdot"] else 4
        return 0.0, tensors * px * a["C"] * es() + (13 + (4 if a["sdot"] else 0) + (4 if a["dq_out"] else 0)) * px, "hbm"
    if name in ("pcm_bn_stats",):
        return 0.0, a["R"] * a["C"] * es(), "hbm"
    if name in ("pcm_bn_apply_fwd",):
        return 0.0, (3 if a["res"] else 2) * a["R"] * a["C"] * es(), "hbm"
    if name == "pcm_bn_bwd_reduce":
        return 0.0, (3 if a["y"] else 2) * a["R"] * a["C"] * es(), "hbm"
    if name == "pcm_bn_bwd_apply":
        return 0.0, (3 + (1 if a["y"] else 0) + (1 if a["dres"] else 0)) * a["R"] * a["C"] * es(), "hbm"
    if name in ("pcm_add", "pcm_relu_bwd"):
        return 0.0, 3 * a["n"] * es(), "hbm"
    if name in ("pcm_dropout", "pcm_add_bcast"):
        return 0.0, 2 * a["n"] * es(), "hbm"
    if name == "pcm_add_layernorm_fwd":
        return 0.0, a["M"] * a["E"] * es() * (2 + (1 if a["b"] else 0) + (1 if a["sum_out"] else 0)), "hbm"
    if name == "pcm_layernorm_bwd":
        return 0.0, 3 * a["M"] * a["E"] * es(), "hbm"
    if name in ("pcm_mha_fwd", "pcm_mha_bwd"):
        E = a["nh"] * a["D"]
        k = 1 if name == "pcm_mha_fwd" else 2.5
        return 4.0 * a["B"] * a["nh"] * a["L"] * a["L"] * a["D"] * k, a["B"] * a["L"] * E * es() * (4 if k == 1 else 8), "tensor"
    if name in ("pcm_pack_weights_batched", "pcm_unpack_grads_batched"):
        return 0.0, 0.0, "hbm"             # sizes live in the device-side job table
    if name == "pcm_conv_gather":
        taps = a["KH"] * a["KW"]
        if a["mode"] == 1 or a["stride"] > 1:
            # strided: every dst pixel sees taps/stride^2 taps (mode 1) or all taps (mode 0)
            taps_eff = taps if a["mode"] == 0 else max(1, taps // (a["stride"] ** 2))
        else:
            taps_eff = taps
        fl = 2.0 * a["N"] * a["Hd"] * a["Wd"] * a["Dc"] * a["Sc"] * taps_eff
        by = a["N"] * (a["Hs"] * a["Ws"] * a["Sc"] * es() + a["Hd"] * a["Wd"] * a["Dc"] * (4 if a["dst_f32"] else es()))
        return fl, by, "tensor"
    if name == "pcm_conv_wgrad":
        fl = 2.0 * a["N"] * a["Ha"] * a["Wa"] * a["Ca"] * a["Cb"] * a["KH"] * a["KW"]
        by = a["N"] * (a["Ha"] * a["Wa"] * a["Ca"] + a["Hb"] * a["Wb"] * a["Cb"]) * es()
        return fl, by, "tensor"
    n_el = None
    if name in ("pcm_gn_stats",):
        return 0.0, a["N"] * a["P"] * a["C"] * es(), "hbm"
    if name in ("pcm_gn_silu_fwd", "pcm_scale_channels"):
        return 0.0, 2 * a["N"] * a["P"] * a["C"] * es(), "hbm"
    if name == "pcm_gn_silu_bwd_reduce":
        return 0.0, 2 * a["N"] * a["P"] * a["C"] * es(), "hbm"                 # da, x in
    if name == "pcm_gn_silu_bwd_apply":
        return 0.0, 3 * a["N"] * a["P"] * a["C"] * es(), "hbm"                 # da, x in; dx out
    if name == "pcm_se_chanstat_fwd":
        return 0.0, a["N"] * a["P"] * (a["C"] * es() + 8), "hbm"               # a in; cmap (2 fp32) out
    if name == "pcm_spatial_gate_fwd":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, px * (2 * a["C"] * es() + 8 + 4), "hbm"                    # a in, out; cmap in, gate out
    if name == "pcm_spatial_gate_bwd_dq":
        return 0.0, a["N"] * a["P"] * (2 * a["C"] * es() + 8), "hbm"           # dout, a in; gate in, dq out
    if name == "pcm_spatial_gate_bwd_dw":
        return 0.0, a["N"] * a["H"] * a["W"] * 12, "hbm"                       # dq, cmap in
    if name == "pcm_spatial_gate_bwd_da":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, px * (3 * a["C"] * es() + 16), "hbm"                       # dout, a in; da out; gate, cmap, dq
    if name == "pcm_maxpool2_fwd":
        return 0.0, a["N"] * a["H"] * a["W"] * a["C"] * es() * 1.25, "hbm"
    if name == "pcm_maxpool2_bwd_skip":
        return 0.0, a["N"] * a["H"] * a["W"] * a["C"] * es() * 2.25, "hbm"     # x in, dx out, dy (1/4) in
    if name == "pcm_maxpool2_bwd_skip_dot":
        px = a["N"] * a["H"] * a["W"]
        return 0.0, px * a["C"] * es() * 2.25 + (4 * px if a["sdot"] else 0), "hbm"
    if name == "pcm_time_mean":
        return 0.0, a["B"] * (a["T"] + 1) * a["P"] * a["C"] * es(), "hbm"
    if name == "pcm_channel_sum":
        return 0.0, a["N"] * a["P"] * a["C"] * es(), "hbm"
    if name in ("pcm_nchw_to_nhwc", "pcm_nhwc_to_nchw"):
        return 0.0, a["N"] * a["H"] * a["W"] * (a["C"] * 4 + a["Cp"] * es()), "hbm"
    if name == "pcm_lstm_cell_fwd":
        return 0.0, a["M"] * a["Ch"] * (16 + 4 + 4 * es() + 4 + es()), "hbm"
    if name == "pcm_lstm_cell_bwd":
        return 0.0, a["M"] * a["Ch"] * (2 * es() + 4 + 4 * es() + 8 + 4 * es() + 4), "hbm"
    if name in ("pcm_head_fwd", "pcm_head_bwd"):
        k = 2 if name == "pcm_head_bwd" else 1
        return 2.0 * a["N"] * a["P"] * a["C"] * a["K"] * k, a["N"] * a["P"] * (a["C"] * es() * k + a["K"] * 4), "hbm"
    if name in ("pcm_head_mse_fwd", "pcm_head_mse_bwd"):
        k = 2 if name == "pcm_head_mse_bwd" else 1                              # x (and dx) + target; pred optional
        return 2.0 * a["N"] * a["P"] * a["C"] * a["K"] * (k + (1 if k == 2 else 0)), a["N"] * a["P"] * (a["C"] * es() * k + a["K"] * 4), "hbm"
    if name in ("pcm_mse_fwd", "pcm_mse_bwd"):
        return 0.0, a["n"] * 4 * (2 if name == "pcm_mse_fwd" else 3), "hbm"
    if name == "pcm_adam_apply":
        return 0.0, a["n"] * 4 * 7, "hbm"
    if name == "pcm_adam_step":
        return 0.0, a["n"] * 4 * 7, "hbm"                                      # p,g,m,v in; p,m,v out
    if name == "pcm_pack_weight":
        return 0.0, a["taps"] * (a["O"] * a["I"] * 4 + a["Op"] * a["Ip"] * es()), "hbm"
    if name == "pcm_metric_partial":
        return 0.0, 2.0 * a["T"] * a["V"] * a["Y"] * a["X"] * 4, "hbm"
    return 0.0, 0.0, "hbm"


def shape_key(name: str, args) -> str:
    """Short signature of the launch's shape (its integer, non-pointer arguments), for grouping identical launches."""
    import ctypes
    ints = [str(x) for (ct, an), x in zip(PROTOS[name][1], args)
            if ct in (ctypes.c_int, ctypes.c_longlong) and an not in ("dtype",)]
    return name[4:] + "[" + ",".join(ints) + "]"
