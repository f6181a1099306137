"""GPU parity cases: run a pcm_b200 module / op on cuda:0 and the oracle on the CPU on identical
seeded inputs, return error metrics.  Used by tests/test_gpu_*.py (asserting) and by
tests/gpu_diag.py (printing a full table without stopping at the first failure)."""
from __future__ import annotations

import math

import numpy as np
import torch

import pcm_b200  # noqa: F401  (registers the package)
from oracle import metric_oracle as MO
from oracle import model_oracle as O
from pcm_b200 import ops
from pcm_b200.config import set_compute_dtype
from tests.golden_util import grad_errors, load_golden, rel_l2

DEV = "cuda:0"


def _oracle_run(fn, sd, x, y, dtype=torch.float64):
    sdd = {k: v.clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    xx = x.clone().to(dtype).requires_grad_(True)
    out = fn(xx, sdd)
    loss = O.mse_loss(out, y.to(dtype))
    loss.backward()
    return out.detach(), float(loss.detach()), {k: v.grad for k, v in sdd.items()}, xx.grad


def _module_run(mod, sd, x, y, need_dx=True):
    mod.load_state_dict(sd, strict=True)
    mod = mod.to(DEV).train()
    xg = x.to(DEV).requires_grad_(need_dx)
    out = mod(xg)
    loss = ops.mse_loss(out, y.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: (p.grad.detach().cpu() if p.grad is not None else None) for k, p in mod.named_parameters()}
    return out.detach().cpu(), float(loss.detach().cpu()), grads, (xg.grad.detach().cpu() if need_dx else None)


def compare(mod, fn, sd, x, y, dtype, golden=None, oracle_keys=None):
    """-> dict(out=rel_l2, loss=rel, dx=rel_l2, grads={name: rel_l2}, golden_* likewise).
    oracle_keys: the entries of sd the oracle function takes (module buffers are left out)."""
    set_compute_dtype(dtype)
    try:
        out, loss, grads, dx = _module_run(mod, sd, x, y)
    finally:
        set_compute_dtype(torch.bfloat16)
    osd = sd if oracle_keys is None else {k: sd[k] for k in oracle_keys}
    o_out, o_loss, o_grads, o_dx = _oracle_run(fn, osd, x, y)
    res = {"out": rel_l2(out.numpy(), o_out.numpy()), "loss": abs(loss - o_loss) / abs(o_loss),
           "dx": rel_l2(dx.numpy(), o_dx.numpy()), "grads": {}, "gnorm": {}}
    for k, og in o_grads.items():
        g = grads.get(k)
        if og is None:
            assert g is None or float(g.abs().max()) == 0.0, f"{k}: oracle grad is None"
            continue
        assert g is not None, f"{k}: missing gradient"
        gn = float(og.norm())
        res["gnorm"][k] = gn
        res["grads"][k] = rel_l2(g.numpy(), og.numpy()) if gn > 1e-7 else float(g.norm())
    if golden is not None:
        z = golden
        res["golden_out"] = rel_l2(out.numpy(), z["out"])
        res["golden_loss"] = abs(loss - float(z["loss"])) / abs(float(z["loss"]))
        ge = grad_errors(grads, z)
        res["golden_grads"] = {k: v[0] for k, v in ge.items()}
        res["golden_gnorm"] = {k: v[2] for k, v in ge.items()}
    return res


# ---- cases ---------------------------------------------------------------------------------------
def case_convblock(dtype):
    from pcm_b200.src.unet import ConvBlock
    cfg, z = load_golden("convblock_small")
    sd = O.synth_state_dict(O._convblock_spec("", cfg["c_in"], cfg["c_out"]), cfg["seed"])
    x, y = O.synth_frame_batch(cfg["B"], cfg["c_in"], cfg["H"], cfg["W"], cfg["seed"] + 1, out_ch=cfg["c_out"])
    return compare(ConvBlock(cfg["c_in"], cfg["c_out"]), lambda a, s: O.conv_block(a, s, ""), sd, x, y, dtype, z)


def case_convblock_ties(dtype):
    """ConvBlock whose output channels 0 and 1 are exact duplicates (same conv2 filters, GroupNorm affine and SE row)
    and biased to be the channel maximum at most pixels: `x.amax(1)` (src/unet.py:27) then splits the gradient evenly
    between the tied channels.  The backward tail relies on the forward tail's saved maximum and tie count and on a
    bit-identical recomputation of a*se; a mismatch would drop the whole max-path gradient."""
    from pcm_b200.src.unet import ConvBlock
    c_in, c_out = 8, 16
    sd = O.synth_state_dict(O._convblock_spec("", c_in, c_out), 41)
    sd["body.3.weight"][1] = sd["body.3.weight"][0]
    sd["body.4.weight"][:2] = 1.0
    sd["body.4.bias"][:2] = 1.5
    sd["se.fc.2.weight"][1] = sd["se.fc.2.weight"][0]
    x, y = O.synth_frame_batch(3, c_in, 12, 20, 42, out_ch=c_out)
    r = compare(ConvBlock(c_in, c_out), lambda a, s: O.conv_block(a, s, ""), sd, x, y, dtype)
    # how often the duplicated pair is the maximum (fp64 oracle): the case is meaningful only if ties are common
    with torch.no_grad():
        a = O.conv_block(x.double(), {k: v.double() for k, v in sd.items()}, "")
    r["tie_share"] = float(((a[:, 0] == a[:, 1]) & (a[:, 0] >= a.amax(1))).double().mean())
    return r


def case_convlstm(dtype):
    from pcm_b200.src.convlstm import ConvLSTM
    cfg, z = load_golden("convlstm_small")
    spec = [("cell.conv.weight", (4 * cfg["c_hid"], cfg["c_in"] + cfg["c_hid"], 3, 3)), ("cell.conv.bias", (4 * cfg["c_hid"],))]
    sd = O.synth_state_dict(spec, cfg["seed"])
    g = torch.Generator().manual_seed(cfg["seed"] + 1)
    x = torch.randn(cfg["T"], cfg["B"], cfg["c_in"], cfg["H"], cfg["W"], generator=g)
    y = torch.randn(cfg["T"], cfg["B"], cfg["c_hid"], cfg["H"], cfg["W"], generator=g)
    return compare(ConvLSTM(cfg["c_in"], cfg["c_hid"]), lambda a, s: O.convlstm(a, s, ""), sd, x, y, dtype, z)


def case_attunet(tag, dtype):
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    cfg, z = load_golden(tag)
    sd = O.synth_state_dict(O.attunet_spec(cfg["in_ch"], cfg["out_ch"], cfg["base"]), cfg["seed"])
    x, y, _ = O.synth_attunet_batch(cfg["B"], cfg["T"], cfg["H"], cfg["W"], cfg["seed"] + 1, cfg["in_ch"], cfg["out_ch"])
    mod = AttUNetConvLSTM(cfg["in_ch"], cfg["out_ch"], cfg["base"], seq_len=cfg["T"])
    return compare(mod, O.attunet_convlstm, sd, x, y, dtype, z)


def case_unet(dtype):
    from pcm_b200.src.unet import UNet
    cfg, z = load_golden("unet_small")
    sd = O.synth_state_dict(O.unet_spec(cfg["in_ch"], cfg["out_ch"], cfg["base"]), cfg["seed"])
    x, y = O.synth_frame_batch(cfg["B"], cfg["in_ch"], cfg["H"], cfg["W"], cfg["seed"] + 1)
    return compare(UNet(cfg["in_ch"], cfg["out_ch"], cfg["base"]), O.unet, sd, x, y, dtype, z)


def case_se(dtype):
    from pcm_b200.src.unet import SEBlock
    sd = O.synth_state_dict([("fc.0.weight", (2, 16, 1, 1)), ("fc.2.weight", (16, 2, 1, 1))], 5)
    x, y = O.synth_frame_batch(3, 16, 10, 14, 6, out_ch=16)
    return compare(SEBlock(16), lambda a, s: O.se_block(a, s, ""), sd, x, y, dtype)


def case_spatial_gate(dtype):
    from pcm_b200.src.unet import SpatialGate
    sd = O.synth_state_dict([("conv.weight", (1, 2, 7, 7))], 7)
    x, y = O.synth_frame_batch(3, 16, 10, 14, 8, out_ch=16)
    return compare(SpatialGate(), lambda a, s: O.spatial_gate(a, s, ""), sd, x, y, dtype)


def case_down_up(dtype):
    """Down then Up with the un-pooled input as skip (odd spatial size exercises floor pooling)."""
    from pcm_b200.src.unet import Down, Up

    class DU(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.down = Down(8, 16)
            self.up = Up(16, 8, 8)

        def forward(self, x):
            return self.up(self.down(x), x)

    spec = O._convblock_spec("down.conv.", 8, 16) + [("up.up.weight", (16, 8, 2, 2)), ("up.up.bias", (8,))] \
        + O._convblock_spec("up.conv.", 16, 8)
    sd = O.synth_state_dict(spec, 9)
    x, y = O.synth_frame_batch(2, 8, 12, 20, 10, out_ch=8)
    fn = lambda a, s: O.up(O.down(a, s, "down."), a, s, "up.")
    return compare(DU(), fn, sd, x, y, dtype)


def case_cell_step(dtype):
    """ConvLSTMCell with explicit non-zero state."""
    from pcm_b200.src.convlstm import ConvLSTMCell
    c_in, c_hid, B, H, W = 8, 8, 2, 6, 9
    sd = O.synth_state_dict([("conv.weight", (4 * c_hid, c_in + c_hid, 3, 3)), ("conv.bias", (4 * c_hid,))], 11)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(B, c_in, H, W, generator=g)
    h0 = torch.randn(B, c_hid, H, W, generator=g) * 0.5
    c0 = torch.randn(B, c_hid, H, W, generator=g)
    yh = torch.randn(B, c_hid, H, W, generator=g)
    yc = torch.randn(B, c_hid, H, W, generator=g)
    set_compute_dtype(dtype)
    try:
        cell = ConvLSTMCell(c_in, c_hid)
        cell.load_state_dict(sd)
        cell = cell.to(DEV)
        xs = [t.to(DEV).requires_grad_(True) for t in (x, h0, c0)]
        h1, c1 = cell(xs[0], (xs[1], xs[2]))
        loss = ops.mse_loss(h1, yh.to(DEV)) + ops.mse_loss(c1, yc.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
    finally:
        set_compute_dtype(torch.bfloat16)
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xo = [t.double().requires_grad_(True) for t in (x, h0, c0)]
    ho, co = O.convlstm_cell(xo[0], xo[1], xo[2], sdd, "")
    lo = O.mse_loss(ho, yh.double()) + O.mse_loss(co, yc.double())
    lo.backward()
    res = {"out": rel_l2(h1.detach().cpu().numpy(), ho.detach().numpy()),
           "loss": abs(float(loss) - float(lo)) / float(lo),
           "dx": rel_l2(xs[0].grad.cpu().numpy(), xo[0].grad.numpy()),
           "grads": {"c_next": rel_l2(c1.detach().cpu().numpy(), co.detach().numpy()),
                     "dh0": rel_l2(xs[1].grad.cpu().numpy(), xo[1].grad.numpy()),
                     "dc0": rel_l2(xs[2].grad.cpu().numpy(), xo[2].grad.numpy())}, "gnorm": {}}
    for k, p in cell.named_parameters():
        res["grads"][k] = rel_l2(p.grad.cpu().numpy(), sdd[k].grad.numpy())
    return res


def case_metric(T=1080):
    from pcm_b200 import metric as M
    pred, true, lat = MO.synth_metric_arrays(T)
    w = MO.get_lat_weights(lat)
    want = {v: MO.metric_triplet(pred[:, i], true[:, i], w) for i, v in enumerate(["tas", "pr"])}
    got = M.weighted_metric_triplets(torch.from_numpy(pred).to(DEV), torch.from_numpy(true).to(DEV), lat)
    err = max(abs(got[i][j] - want[v][j]) / abs(want[v][j]) for i, v in enumerate(["tas", "pr"]) for j in range(3))
    sc = M.combined_score({v: got[i] for i, v in enumerate(["tas", "pr"])})
    return {"max_rel": err, "score_rel": abs(sc - MO.combined_score(want)) / MO.combined_score(want)}


def case_metric_appendix_g():
    from pcm_b200 import metric as M
    fx = MO.known_answer_fixture()
    pred = np.stack([fx["tas_pred"], fx["pr_pred"]], 1).astype(np.float32)
    true = np.stack([fx["tas_true"], fx["pr_true"]], 1).astype(np.float32)
    got = M.weighted_metric_triplets(torch.from_numpy(pred).to(DEV), torch.from_numpy(true).to(DEV), fx["lats"])
    want = {"tas": [1.9653055953, 0.5983730314, 0.5430399866], "pr": [0.9742019753, 0.3126719193, 0.7152702066]}
    err = max(abs(got[i][j] - want[v][j]) / want[v][j] for i, v in enumerate(["tas", "pr"]) for j in range(3))
    sc = M.combined_score({v: got[i] for i, v in enumerate(["tas", "pr"])})
    return {"max_rel": err, "score_rel": abs(sc - 1.1422441747) / 1.1422441747}


def case_adam():
    from pcm_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    ps = [torch.randn(1000, generator=g), torch.randn(37, 5, generator=g)]
    params = [torch.nn.Parameter(p.clone().to(DEV)) for p in ps]
    opt = FusedAdam(params, lr=5e-4)
    ref = [p.clone() for p in ps]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for step in range(1, 5):
        gr = [torch.randn(p.shape, generator=g) for p in ps]
        for p, gg in zip(params, gr):
            p.main_grad.copy_(gg.to(DEV))
        opt.step()
        for p, gg, m, v in zip(ref, gr, ms, vs):
            O.adam_step(p, gg, m, v, step)
    torch.cuda.synchronize()
    return {"max_rel": max(rel_l2(p.detach().cpu().numpy(), r.numpy()) for p, r in zip(params, ref))}


def case_season_stage():
    g = torch.Generator().manual_seed(4)
    x5 = torch.randn(6, 5, 8, 12, generator=g)
    month = torch.randint(0, 12, (6,), generator=g)
    y = ops.season_embed_stage(x5.to(DEV), month.to(DEV), torch.float32).cpu()
    ang = 2 * math.pi * month.float() / 12
    want = torch.zeros(6, 8, 12, 16)
    want[..., :5] = x5.permute(0, 2, 3, 1)
    want[..., 5] = torch.sin(ang)[:, None, None]
    want[..., 6] = torch.cos(ang)[:, None, None]
    return {"max_abs": float((y - want).abs().max())}


def _buffers_state(sd, mod):
    """oracle state dicts hold parameters only; BN buffers keep the constructor defaults."""
    full = dict(mod.state_dict())
    full.update(sd)
    return full


def case_simplecnn(dtype, tag="simplecnn_small"):
    from pcm_b200.src.models import SimpleCNN
    cfg, z = load_golden(tag)
    sd = O.synth_state_dict(O.simplecnn_spec(cfg["n_in"], cfg["n_out"], 3, cfg["init_dim"], cfg["depth"]), cfg["seed"])
    x, y = O.synth_frame_batch(cfg["B"], cfg["n_in"], cfg["H"], cfg["W"], cfg["seed"] + 1)
    mod = SimpleCNN(cfg["n_in"], cfg["n_out"], kernel_size=3, init_dim=cfg["init_dim"], depth=cfg["depth"], dropout_rate=0.0)
    full = _buffers_state(sd, mod)
    res = compare(mod, lambda a, s: O.simple_cnn(a, s, cfg["depth"]), full, x, y, dtype, z,
                  oracle_keys=list(sd.keys()))
    # running statistics after one training step (momentum 0.1, unbiased variance) vs torch's own BatchNorm
    with torch.no_grad():
        y0 = torch.nn.functional.conv2d(x.double(), sd["initial.0.weight"].double(), sd["initial.0.bias"].double(), padding=1)
        rm = 0.1 * y0.mean(dim=(0, 2, 3))
        rv = 0.9 + 0.1 * y0.var(dim=(0, 2, 3), unbiased=True)
    res["running_mean"] = rel_l2(mod.initial[1].running_mean.cpu().numpy(), rm.numpy())
    res["running_var"] = rel_l2(mod.initial[1].running_var.cpu().numpy(), rv.numpy())
    res["nbt"] = int(mod.initial[1].num_batches_tracked)
    return res


def case_cnn_transformer(dtype, tag="cnn_transformer_small"):
    from pcm_b200.src.cnn_transformer import CNNTransformer
    cfg, z = load_golden(tag)
    sd = O.synth_state_dict(O.cnn_transformer_spec(5, 2, cfg["embed_dim"], cfg["depth"], cfg["n_heads"], cfg["mlp_dim"]),
                            cfg["seed"])
    x, y = O.synth_frame_batch(cfg["B"], 5, 48, 72, cfg["seed"] + 1)
    mod = CNNTransformer(5, 2, cfg["embed_dim"], cfg["depth"], cfg["n_heads"], cfg["mlp_dim"], dropout=0.0)
    return compare(mod, lambda a, s: O.cnn_transformer(a, s, cfg["depth"], cfg["n_heads"]), sd, x, y, dtype, z)


def case_full_size(kind, dtype=torch.bfloat16):
    """Default-size models (BASELINE configs[0] / configs[1] geometry, reduced batch) on the tensor-core paths
    against the fp64 oracle."""
    if kind == "simplecnn":
        from pcm_b200.src.models import SimpleCNN
        sd = O.synth_state_dict(O.simplecnn_spec(5, 2, 3, 64, 4), 171)
        x, y = O.synth_frame_batch(2, 5, 48, 72, 172)
        mod = SimpleCNN(5, 2, dropout_rate=0.0)
        return compare(mod, lambda a, s: O.simple_cnn(a, s, 4), _buffers_state(sd, mod), x, y, dtype,
                       oracle_keys=list(sd.keys()))
    from pcm_b200.src.cnn_transformer import CNNTransformer
    sd = O.synth_state_dict(O.cnn_transformer_spec(5, 2, 128, 4, 4, 256), 181)
    x, y = O.synth_frame_batch(4, 5, 48, 72, 182)
    mod = CNNTransformer(dropout=0.0)
    return compare(mod, lambda a, s: O.cnn_transformer(a, s, 4, 4), sd, x, y, dtype)


def case_dropout_stats():
    """Counter-based dropout: keep rate, 1/(1-p) scaling, backward uses the same mask, Dropout2d drops whole channels."""
    from pcm_b200 import ops_nn
    p = 0.2
    x = torch.ones(8, 12, 18, 64, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    y = ops_nn.DropoutFn.apply(x, p, 12345)
    y.sum().backward()
    keep = float((y != 0).float().mean())
    same = bool(((x.grad != 0) == (y != 0)).all())
    scale = float(y.max())
    x2 = torch.ones(16, 6, 9, 64, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    y2 = ops_nn.Dropout2dFn.apply(x2, p, 999)
    per_ch = y2.float().reshape(16, 54, 64)
    whole = bool(((per_ch == per_ch[:, :1, :]).all()))
    keep2 = float((per_ch[:, 0, :] != 0).float().mean())
    return {"keep": keep, "same_mask": same, "scale": scale, "whole_channels": whole, "keep2d": keep2}


def case_config5_geometry():
    """BASELINE configs[4] geometry (1-degree grid 180x360 zero-padded to 184x360, base channels x2), tiny batch: the
    tensor-core convolutions take the large grid (TMA boxes, tile chooser), the ConvBlock tails fall back to the
    grid-wide kernels (an image no longer fits an SM), the weight gradients of the 360-wide levels to the SIMT kernel.
    Self-consistency against the fp64 oracle on the same inputs."""
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    B, T, H, W, base = 1, 2, 184, 360, 32
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 191)
    x, y, _ = O.synth_attunet_batch(B, T, H, W, 192)
    mod = AttUNetConvLSTM(7, 2, base, seq_len=T)
    return compare(mod, O.attunet_convlstm, sd, x, y, torch.bfloat16)


def case_simplecnn_eval():
    """model.eval(): BatchNorm uses the running statistics, Dropout2d is the identity (src/models.py:114-123 under
    Lightning's validation loop) — against torch's functional batch_norm with the same buffers."""
    import torch.nn.functional as F
    from pcm_b200.src.models import SimpleCNN
    sd = O.synth_state_dict(O.simplecnn_spec(5, 2, 3, 16, 3), 201)
    mod = SimpleCNN(5, 2, kernel_size=3, init_dim=16, depth=3, dropout_rate=0.2)
    full = _buffers_state(sd, mod)
    g = torch.Generator().manual_seed(202)
    for k in list(full.keys()):
        if k.endswith("running_mean"):
            full[k] = torch.randn(full[k].shape, generator=g) * 0.3
        elif k.endswith("running_var"):
            full[k] = torch.rand(full[k].shape, generator=g) + 0.5
    mod.load_state_dict(full)
    mod = mod.to(DEV).eval()
    x, _ = O.synth_frame_batch(3, 5, 16, 24, 203)
    set_compute_dtype(torch.float32)
    try:
        with torch.no_grad():
            out = mod(x.to(DEV)).cpu()
    finally:
        set_compute_dtype(torch.bfloat16)

    def bn(t, p):
        return F.batch_norm(t, full[p + "running_mean"].double(), full[p + "running_var"].double(),
                            full[p + "weight"].double(), full[p + "bias"].double(), False, 0.1, 1e-5)

    def conv(t, p, pad):
        return F.conv2d(t, full[p + "weight"].double(), full[p + "bias"].double(), padding=pad)

    t = F.relu(bn(conv(x.double(), "initial.0.", 1), "initial.1."))
    for i in range(3):
        p = f"res_blocks.{i}."
        o = F.relu(bn(conv(t, p + "conv1.", 1), p + "bn1."))
        o = bn(conv(o, p + "conv2.", 1), p + "bn2.")
        idt = bn(conv(t, p + "skip.0.", 0), p + "skip.1.") if (p + "skip.0.weight") in full else t
        t = F.relu(o + idt)
    t = F.relu(bn(conv(t, "final.0.", 1), "final.1."))
    want = conv(t, "final.3.", 0)
    return {"out": rel_l2(out.numpy(), want.numpy())}


def case_attunet_b64(tag, dtype, through_trainer=True):
    """The BENCHMARK shape (B=64, T=6, 48x72, base 16) against the fixture the real reference produced
    (tests/golden/<tag>.npz: strided samples of the output and of every gradient, loss).  The step runs through
    TrainStep under the captured CUDA graph with the second stream, i.e. exactly what bench.py times: the loss and the
    gradients left in the flat buffer by the replayed step are compared; the output comes from an eager forward.
    tag 'attunet_cfg3_b64': seed-derived weights; 'attunet_cfg3_b64_default_init': the reference's own default
    initialisation under torch.manual_seed(42) on bench.py's first batch."""
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    cfg, z = load_golden(tag)
    B, T, H, W = cfg["B"], cfg["T"], cfg["H"], cfg["W"]
    if cfg["init"] == "synth":
        mod = AttUNetConvLSTM(cfg["in_ch"], cfg["out_ch"], cfg["base"], seq_len=T)
        mod.load_state_dict(O.synth_state_dict(O.attunet_spec(cfg["in_ch"], cfg["out_ch"], cfg["base"]), cfg["seed"]))
        x, y, _ = O.synth_attunet_batch(B, T, H, W, cfg["seed"] + 1, cfg["in_ch"], cfg["out_ch"])
    else:
        torch.manual_seed(cfg["seed"])
        mod = AttUNetConvLSTM(cfg["in_ch"], cfg["out_ch"], cfg["base"], seq_len=T)
        x, y, _ = O.synth_attunet_batch(B, T, H, W, cfg["batch_seed"], cfg["in_ch"], cfg["out_ch"])
    set_compute_dtype(dtype)
    try:
        mod = mod.to(DEV).train()
        xd, yd = x.to(DEV), y.to(DEV)
        with torch.no_grad():
            out = mod(xd).float().cpu().numpy().reshape(-1)
        step = TrainStep(mod, tuple(x.shape), tuple(y.shape), lr=5e-4, use_graph=through_trainer)
        step.load_batch(xd, yd)
        if through_trainer:
            step.warmup_and_capture(warmup=2)            # restores the parameters afterwards
            assert step.graph is not None and step.side is not None
        loss = float(step.step(xd, yd).item())
        torch.cuda.synchronize()
        grads = {k: p.main_grad.detach().cpu().clone() for k, p in mod.named_parameters() if not k.startswith("post_conv")}
        for k, p in mod.named_parameters():
            if k.startswith("post_conv"):
                grads[k] = None if float(p.main_grad.abs().max()) == 0.0 else p.main_grad.detach().cpu()
    finally:
        set_compute_dtype(torch.bfloat16)
    stride = max(1, out.size // (4 * 1024))
    ge = grad_errors(grads, z)
    return {"out": rel_l2(out[::stride], z["out_sample"]),
            "out_norm": abs(float(np.linalg.norm(out.astype(np.float64))) - float(z["out_norm"][0])) / float(z["out_norm"][0]),
            "loss": abs(loss - float(z["loss"])) / abs(float(z["loss"])),
            "grads": {k: v[0] for k, v in ge.items()}, "gnorm": {k: v[2] for k, v in ge.items()}}
