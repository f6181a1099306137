"""CPU oracle for the emulator's model hot path (TEST INFRASTRUCTURE — not product code).

A functional restatement, in plain torch CPU ops over a flat ``{state_dict key: tensor}``
mapping, of the reference's forward passes.  Gradients come from torch CPU autograd over
these functions.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product path
(``physics-based-climate-model_b200/``) never does.

Pinning: the reference holds NO golden vectors for the models.  The
restatement is pinned against the *imported reference modules themselves* in the build
container (``oracle/make_goldens.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures on every CPU run, and ``tests/test_cpu_boundary.py`` re-checks the parameter
surface / default initialisation live against the reference modules whenever ``/root/reference`` is present.

Every function cites the reference lines (relative to /root/reference) it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

GN_GROUPS = 8      # src/unet.py:37,39  nn.GroupNorm(8, c_out)
GN_EPS = 1e-5      # torch default, src/unet.py:37
SE_RATIO = 8       # src/unet.py:8


# --------------------------------------------------------------------------------------
# parameter surfaces — name -> shape, in state_dict order
# --------------------------------------------------------------------------------------
def _convblock_spec(p: str, c_in: int, c_out: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Keys of one ConvBlock (src/unet.py:32-42)."""
    return [
        (p + "body.0.weight", (c_out, c_in, 3, 3)),
        (p + "body.1.weight", (c_out,)),
        (p + "body.1.bias", (c_out,)),
        (p + "body.3.weight", (c_out, c_out, 3, 3)),
        (p + "body.4.weight", (c_out,)),
        (p + "body.4.bias", (c_out,)),
        (p + "se.fc.0.weight", (c_out // SE_RATIO, c_out, 1, 1)),
        (p + "se.fc.2.weight", (c_out, c_out // SE_RATIO, 1, 1)),
        (p + "spat.conv.weight", (1, 2, 7, 7)),
    ]


def attunet_spec(in_ch: int = 7, out_ch: int = 2, base: int = 16):
    """src/unet_convlstm_attention.py:28-56 — registration order of the reference ctor."""
    b = base
    s = []
    s += _convblock_spec("enc1.", in_ch, b)
    s += _convblock_spec("enc2.conv.", b, 2 * b)
    s += _convblock_spec("enc3.conv.", 2 * b, 4 * b)
    s += _convblock_spec("enc4.conv.", 4 * b, 8 * b)
    s += [("convlstm.cell.conv.weight", (16 * b, 12 * b, 3, 3)), ("convlstm.cell.conv.bias", (16 * b,))]
    s += [("post_conv.0.weight", (4 * b, 4 * b, 3, 3)), ("post_conv.0.bias", (4 * b,))]   # dead, :46-49
    s += [("up3.up.weight", (4 * b, 4 * b, 2, 2)), ("up3.up.bias", (4 * b,))]
    s += _convblock_spec("up3.conv.", 8 * b, 4 * b)
    s += [("up2.up.weight", (4 * b, 2 * b, 2, 2)), ("up2.up.bias", (2 * b,))]
    s += _convblock_spec("up2.conv.", 4 * b, 2 * b)
    s += [("up1.up.weight", (2 * b, b, 2, 2)), ("up1.up.bias", (b,))]
    s += _convblock_spec("up1.conv.", 2 * b, b)
    s += [("head.weight", (out_ch, b, 1, 1)), ("head.bias", (out_ch,))]
    return s


def unet_spec(in_ch: int = 5, out_ch: int = 2, base: int = 16):
    """src/unet.py:78-96."""
    b = base
    s = []
    s += _convblock_spec("enc1.", in_ch, b)
    s += _convblock_spec("enc2.conv.", b, 2 * b)
    s += _convblock_spec("enc3.conv.", 2 * b, 4 * b)
    s += _convblock_spec("enc4.conv.", 4 * b, 8 * b)
    s += _convblock_spec("bott.", 8 * b, 8 * b)
    s += [("up3.up.weight", (8 * b, 4 * b, 2, 2)), ("up3.up.bias", (4 * b,))]
    s += _convblock_spec("up3.conv.", 8 * b, 4 * b)
    s += [("up2.up.weight", (4 * b, 2 * b, 2, 2)), ("up2.up.bias", (2 * b,))]
    s += _convblock_spec("up2.conv.", 4 * b, 2 * b)
    s += [("up1.up.weight", (2 * b, b, 2, 2)), ("up1.up.bias", (b,))]
    s += _convblock_spec("up1.conv.", 2 * b, b)
    s += [("head.weight", (out_ch, b, 1, 1)), ("head.bias", (out_ch,))]
    return s


def cnn_transformer_spec(in_channels=5, out_channels=2, embed_dim=128, depth=4, n_heads=4, mlp_dim=256):
    """src/cnn_transformer.py:5-41."""
    e = embed_dim
    s = [("pos_embedding", (1, 216, e)),
         ("encoder.0.weight", (e // 2, in_channels, 3, 3)), ("encoder.0.bias", (e // 2,)),
         ("encoder.2.weight", (e, e // 2, 3, 3)), ("encoder.2.bias", (e,))]
    for l in range(depth):
        p = f"transformer.layers.{l}."
        s += [(p + "self_attn.in_proj_weight", (3 * e, e)), (p + "self_attn.in_proj_bias", (3 * e,)),
              (p + "self_attn.out_proj.weight", (e, e)), (p + "self_attn.out_proj.bias", (e,)),
              (p + "linear1.weight", (mlp_dim, e)), (p + "linear1.bias", (mlp_dim,)),
              (p + "linear2.weight", (e, mlp_dim)), (p + "linear2.bias", (e,)),
              (p + "norm1.weight", (e,)), (p + "norm1.bias", (e,)),
              (p + "norm2.weight", (e,)), (p + "norm2.bias", (e,))]
    s += [("decoder.0.weight", (e, e // 2, 2, 2)), ("decoder.0.bias", (e // 2,)),
          ("decoder.2.weight", (e // 2, e // 4, 2, 2)), ("decoder.2.bias", (e // 4,)),
          ("decoder.4.weight", (out_channels, e // 4, 1, 1)), ("decoder.4.bias", (out_channels,))]
    return s


def simplecnn_spec(n_in=5, n_out=2, kernel_size=3, init_dim=64, depth=4):
    """src/models.py:76-112 (parameters only; BN buffers are handled by ``simplecnn_buffers``)."""
    k = kernel_size
    s = [("initial.0.weight", (init_dim, n_in, k, k)), ("initial.0.bias", (init_dim,)),
         ("initial.1.weight", (init_dim,)), ("initial.1.bias", (init_dim,))]
    cur = init_dim
    for i in range(depth):
        out = cur * 2 if i < depth - 1 else cur
        p = f"res_blocks.{i}."
        s += [(p + "conv1.weight", (out, cur, k, k)), (p + "conv1.bias", (out,)),
              (p + "bn1.weight", (out,)), (p + "bn1.bias", (out,)),
              (p + "conv2.weight", (out, out, k, k)), (p + "conv2.bias", (out,)),
              (p + "bn2.weight", (out,)), (p + "bn2.bias", (out,))]
        if cur != out:
            s += [(p + "skip.0.weight", (out, cur, 1, 1)), (p + "skip.0.bias", (out,)),
                  (p + "skip.1.weight", (out,)), (p + "skip.1.bias", (out,))]
        cur = out
    s += [("final.0.weight", (cur // 2, cur, k, k)), ("final.0.bias", (cur // 2,)),
          ("final.1.weight", (cur // 2,)), ("final.1.bias", (cur // 2,)),
          ("final.3.weight", (n_out, cur // 2, 1, 1)), ("final.3.bias", (n_out,))]
    return s


def synth_state_dict(spec, seed: int, dtype=torch.float32) -> SD:
    """Deterministic synthetic weights (not the reference init — a recipe both the oracle, the
    real reference (in make_goldens.py) and the CUDA path can regenerate from a seed alone).
    conv/linear weights ~ N(0, 1/fan_in); norm weights ~ 1 + 0.1 N; biases ~ 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in spec:
        if len(shape) == 1:
            base = 1.0 if (name.endswith("weight")) else 0.0
            t = base + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif name == "pos_embedding":
            t = torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            if ".up.weight" in name or name.startswith("decoder.0") or name.startswith("decoder.2"):
                fan_in = shape[0]          # ConvTranspose2d layout is (C_in, C_out, kH, kW)
            t = torch.randn(shape, generator=g, dtype=torch.float64) * (1.5 / math.sqrt(fan_in))
        sd[name] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def se_block(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/unet.py:16-17 — x * sigmoid(W2 relu(W1 avgpool(x))), both 1x1 convs bias-free."""
    s = x.mean(dim=(2, 3), keepdim=True)
    s = F.relu(F.conv2d(s, sd[p + "fc.0.weight"]))
    s = torch.sigmoid(F.conv2d(s, sd[p + "fc.2.weight"]))
    return x * s


def spatial_gate(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/unet.py:25-29 — x * sigmoid(conv7x7([mean_C x, amax_C x]))."""
    m = torch.cat([x.mean(1, keepdim=True), x.amax(1, keepdim=True)], dim=1)
    return x * torch.sigmoid(F.conv2d(m, sd[p + "conv.weight"], padding=3))


def conv_block(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/unet.py:44-49 with body :35-40 — [conv3x3 -> GroupNorm(8) -> SiLU] x2 -> SE -> SpatialGate."""
    y = F.conv2d(x, sd[p + "body.0.weight"], padding=1)
    y = F.silu(F.group_norm(y, GN_GROUPS, sd[p + "body.1.weight"], sd[p + "body.1.bias"], GN_EPS))
    y = F.conv2d(y, sd[p + "body.3.weight"], padding=1)
    y = F.silu(F.group_norm(y, GN_GROUPS, sd[p + "body.4.weight"], sd[p + "body.4.bias"], GN_EPS))
    y = se_block(y, sd, p + "se.")
    return spatial_gate(y, sd, p + "spat.")


def down(x: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/unet.py:57-58 / src/unet_convlstm_attention.py:24-25 — MaxPool2d(2) (floor) -> ConvBlock."""
    return conv_block(F.max_pool2d(x, 2), sd, p + "conv.")


def up(x: torch.Tensor, skip: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/unet.py:66-69 — ConvTranspose2d(k2,s2,bias) -> cat([up, skip]) -> ConvBlock."""
    x = F.conv_transpose2d(x, sd[p + "up.weight"], sd[p + "up.bias"], stride=2)
    return conv_block(torch.cat([x, skip], dim=1), sd, p + "conv.")


def convlstm_cell(x, h, c, sd: SD, p: str):
    """src/convlstm.py:11-19 — gates = conv(cat[x,h]); order i,f,o,g."""
    w = sd[p + "conv.weight"]
    gates = F.conv2d(torch.cat([x, h], dim=1), w, sd[p + "conv.bias"], padding=w.shape[-1] // 2)
    i, f, o, g = gates.chunk(4, dim=1)
    i, f, o, g = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(g)
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def convlstm(x_seq: torch.Tensor, sd: SD, p: str) -> torch.Tensor:
    """src/convlstm.py:27-35 — x_seq (T,B,C,H,W) -> stacked h (T,B,C_hid,H,W), zero init."""
    c_hid = sd[p + "cell.conv.weight"].shape[0] // 4
    T, B, _, H, W = x_seq.shape
    h = x_seq.new_zeros(B, c_hid, H, W)
    c = torch.zeros_like(h)
    outs = []
    for t in range(T):
        h, c = convlstm_cell(x_seq[t], h, c, sd, p + "cell.")
        outs.append(h)
    return torch.stack(outs)


# --------------------------------------------------------------------------------------
# models
# --------------------------------------------------------------------------------------
def attunet_convlstm(x_seq: torch.Tensor, sd: SD) -> torch.Tensor:
    """src/unet_convlstm_attention.py:60-104 — x_seq (B,T,C,H,W) -> (B,out_ch,H,W).
    The per-frame encoder is frame-independent (GroupNorm is per sample), so T is folded
    into the batch here; only the ConvLSTM is sequential."""
    B, T, C, H, W = x_seq.shape
    x = x_seq.reshape(B * T, C, H, W)
    s1 = conv_block(x, sd, "enc1.")
    s2 = down(s1, sd, "enc2.")
    s3 = down(s2, sd, "enc3.")
    s4 = down(s3, sd, "enc4.")
    lstm_in = s4.reshape(B, T, *s4.shape[1:]).transpose(0, 1)          # :85 stack over t
    bott = convlstm(lstm_in, sd, "convlstm.")[-1]                       # :87-88
    sk = lambda s: s.reshape(B, T, *s.shape[1:]).mean(dim=1)            # :91-93 time-mean skips
    d3 = up(bott, sk(s3), sd, "up3.")
    d2 = up(d3, sk(s2), sd, "up2.")
    d1 = up(d2, sk(s1), sd, "up1.")
    return F.conv2d(d1, sd["head.weight"], sd["head.bias"])             # :104


def unet(x: torch.Tensor, sd: SD) -> torch.Tensor:
    """src/unet.py:98-109."""
    s1 = conv_block(x, sd, "enc1.")
    s2 = down(s1, sd, "enc2.")
    s3 = down(s2, sd, "enc3.")
    s4 = down(s3, sd, "enc4.")
    y = conv_block(s4, sd, "bott.")
    y = up(y, s3, sd, "up3.")
    y = up(y, s2, sd, "up2.")
    y = up(y, s1, sd, "up1.")
    return F.conv2d(y, sd["head.weight"], sd["head.bias"])


def transformer_layer(x: torch.Tensor, sd: SD, p: str, n_heads: int) -> torch.Tensor:
    """One post-norm nn.TransformerEncoderLayer (src/cnn_transformer.py:25-31), dropout = 0:
    LN1(x + out_proj(softmax(q k^T / sqrt(d)) v)) -> LN2(y + W2 relu(W1 y))."""
    B, L, E = x.shape
    d = E // n_heads
    qkv = F.linear(x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(E, dim=-1)
    sh = lambda t: t.reshape(B, L, n_heads, d).transpose(1, 2)
    att = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / math.sqrt(d), dim=-1) @ sh(v)
    att = att.transpose(1, 2).reshape(B, L, E)
    att = F.linear(att, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    y = F.layer_norm(x + att, (E,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    ff = F.linear(F.relu(F.linear(y, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                  sd[p + "linear2.weight"], sd[p + "linear2.bias"])
    return F.layer_norm(y + ff, (E,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)


def cnn_transformer(x: torch.Tensor, sd: SD, depth: int = 4, n_heads: int = 4) -> torch.Tensor:
    """src/cnn_transformer.py:43-54 with dropout disabled (p=0 / eval)."""
    B = x.shape[0]
    y = F.relu(F.conv2d(x, sd["encoder.0.weight"], sd["encoder.0.bias"], stride=2, padding=1))
    y = F.relu(F.conv2d(y, sd["encoder.2.weight"], sd["encoder.2.bias"], stride=2, padding=1))
    E, Hh, Ww = y.shape[1:]
    y = y.flatten(2).transpose(1, 2) + sd["pos_embedding"]
    for l in range(depth):
        y = transformer_layer(y, sd, f"transformer.layers.{l}.", n_heads)
    y = y.transpose(1, 2).reshape(B, E, Hh, Ww)
    y = F.relu(F.conv_transpose2d(y, sd["decoder.0.weight"], sd["decoder.0.bias"], stride=2))
    y = F.relu(F.conv_transpose2d(y, sd["decoder.2.weight"], sd["decoder.2.bias"], stride=2))
    return F.conv2d(y, sd["decoder.4.weight"], sd["decoder.4.bias"])


def _bn_train(x, sd: SD, p: str):
    """nn.BatchNorm2d in training mode (batch statistics, biased var, eps 1e-5) — src/models.py:48,51,57,91,109."""
    return F.batch_norm(x, None, None, sd[p + "weight"], sd[p + "bias"], True, 0.1, 1e-5)


def residual_block(x, sd: SD, p: str):
    """src/models.py:60-73."""
    k = sd[p + "conv1.weight"].shape[-1]
    out = F.relu(_bn_train(F.conv2d(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=k // 2), sd, p + "bn1."))
    out = _bn_train(F.conv2d(out, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=k // 2), sd, p + "bn2.")
    if (p + "skip.0.weight") in sd:
        idt = _bn_train(F.conv2d(x, sd[p + "skip.0.weight"], sd[p + "skip.0.bias"]), sd, p + "skip.1.")
    else:
        idt = x
    return F.relu(out + idt)


def simple_cnn(x, sd: SD, depth: int = 4):
    """src/models.py:114-123, training-mode BN, Dropout2d disabled (p=0)."""
    k = sd["initial.0.weight"].shape[-1]
    y = F.relu(_bn_train(F.conv2d(x, sd["initial.0.weight"], sd["initial.0.bias"], padding=k // 2), sd, "initial.1."))
    for i in range(depth):
        y = residual_block(y, sd, f"res_blocks.{i}.")
    y = F.relu(_bn_train(F.conv2d(y, sd["final.0.weight"], sd["final.0.bias"], padding=k // 2), sd, "final.1."))
    return F.conv2d(y, sd["final.3.weight"], sd["final.3.bias"])


def mse_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """main_final.py:544,559 — nn.MSELoss() (mean over all elements)."""
    return ((pred - target) ** 2).mean()


def adam_step(p, g, m, v, step: int, lr=5e-4, b1=0.9, b2=0.999, eps=1e-8, wd=0.0):
    """torch.optim.Adam as configured at main_final.py:742-746 (L2-style weight decay, no amsgrad).
    In-place on p, m, v; ``step`` is 1-based."""
    if wd != 0.0:
        g = g + wd * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synth_attunet_batch(B: int, T: int, H: int, W: int, seed: int = 42, in_ch: int = 7, out_ch: int = 2):
    """Config-3 inputs: ch 0-4 ~ N(0,1); ch 5/6 = sin/cos(2 pi m/12) broadcast (main_final.py:188-196);
    1/16 of samples get k~U{0..T-1} leading zero frames (left pad of main_final.py:127-131).
    Returns x (B,T,in_ch,H,W) fp32, y (B,out_ch,H,W) fp32, month0 (B,) int64."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, in_ch, H, W, generator=g)
    y = torch.randn(B, out_ch, H, W, generator=g)
    m0 = torch.randint(0, 12, (B,), generator=g)
    if in_ch >= 7:
        m = (m0[:, None] + torch.arange(T)[None, :]) % 12
        ang = 2 * math.pi * m.to(torch.float32) / 12
        x[:, :, 5] = torch.sin(ang)[:, :, None, None]
        x[:, :, 6] = torch.cos(ang)[:, :, None, None]
    npad = max(1, B // 16) if B >= 2 else 0
    if npad and T > 1:
        k = torch.randint(0, T, (npad,), generator=g)
        for j in range(npad):
            x[j * (B // npad), : int(k[j])] = 0.0
    return x, y, m0


def synth_frame_batch(B: int, C: int, H: int, W: int, seed: int = 42, out_ch: int = 2):
    """Configs 1/2: single-frame x ~ N(0,1) (B,C,H,W), y ~ N(0,1) (B,out_ch,H,W)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, C, H, W, generator=g), torch.randn(B, out_ch, H, W, generator=g)
