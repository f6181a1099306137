"""GPU parity tests (run with -m gpu on a B200): the pcm_b200 CUDA path, called through the C ABI,
against the CPU oracle on identical seeded inputs and against the golden fixtures produced by the
real reference modules.

Tolerances (stated from data — DESIGN.md §parity):
  fp32 compute path : out rel-L2 <= 1e-4, loss <= 1e-5, every gradient rel-L2 <= 2e-3
                      (atomics reorder fp32 sums; observed <= 4e-4)
  bf16 compute path : out rel-L2 <= 3e-2, loss <= 2e-3, gradients: median rel-L2 <= 0.12, every tensor
                      <= 0.5 except the squeeze-excitation bottleneck weights (2-16 ReLU units whose
                      on/off state flips under bf16 rounding; the reference's own bf16-autocast run
                      shows 0.5 there) which must stay within 4x in norm.
  For calibration the reference under torch.autocast(bf16) vs its fp32 self on these same cases gives
  out 1.5e-2..2.6e-2, gradient median 0.07..0.10, worst 0.24..0.50 (measured in the build container).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# SimpleCNN stacks ten BatchNorm layers, which amplify rounding: the REFERENCE under torch.autocast(bf16) against its
# own fp32 run gives a gradient median rel-L2 of 0.18 (small case) / 0.20 (default widths), worst tensor 0.34 —
# measured in the build container on the same seeds (see DESIGN.md parity section); ours must stay within that.
SIMPLECNN_BF16_MEDIAN = 0.25
F32 = dict(out=1e-4, loss=1e-5, grad=2e-3)
BF16 = dict(out=3e-2, loss=2e-3, grad=0.5, median=0.12)


def _check(res, dtype, median=None):
    tol = dict(F32 if dtype == torch.float32 else BF16)
    if median is not None:
        tol["median"] = median
    assert res["out"] < tol["out"], res["out"]
    assert res["loss"] < tol["loss"], res["loss"]
    if "golden_out" in res:
        # the same run against the fixtures the REAL reference modules produced (oracle/make_goldens.py)
        assert res["golden_out"] < tol["out"], res["golden_out"]
        assert res["golden_loss"] < tol["loss"], res["golden_loss"]
        gerrs = []
        for k, e in res["golden_grads"].items():
            if res["golden_gnorm"][k] <= 1e-7:
                assert e < (1e-4 if dtype == torch.float32 else 1e-3), (k, e)
                continue
            gerrs.append(e)
            if dtype != torch.float32 and ".se.fc." in k or k.startswith("fc."):
                assert e < 4.0, (k, e)
            else:
                assert e < tol["grad"], ("golden", k, e)
        if dtype != torch.float32 and len(gerrs) >= 8:
            assert float(np.median(gerrs)) < tol["median"], float(np.median(gerrs))
    errs = []
    for k, e in res["grads"].items():
        gn = res["gnorm"].get(k, 1.0)
        if gn <= 1e-7:
            # exactly-zero gradient in the oracle (dead unit, or a conv bias cancelled by the BatchNorm that follows):
            # ours must be ~0 too, i.e. no larger than the rounding residue of summing O(10^3..10^5) terms of
            # magnitude ~1e-3 whose exact sum is zero (fp32 atomics in varying order: up to ~1e-5; bf16-stored
            # terms: ~1e-4)
            assert e < (1e-4 if dtype == torch.float32 else 1e-3), (k, e)
            continue
        errs.append(e)
        if dtype != torch.float32 and ".se.fc." in k or k.startswith("fc."):
            assert e < 4.0, (k, e)
        else:
            assert e < tol["grad"], (k, e, gn)
    if dtype != torch.float32 and len(errs) >= 8:
        assert float(np.median(errs)) < tol["median"], float(np.median(errs))
    assert res["dx"] < (tol["grad"] if dtype == torch.float32 else 0.5), res["dx"]


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from tests import gpu_cases
    return gpu_cases


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", ["se", "spatial_gate", "convblock", "down_up", "cell_step", "convlstm", "unet"])
def test_blocks(G, case, dtype):
    _check(getattr(G, "case_" + case)(dtype), dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_convblock_tied_channel_maximum(G, dtype):
    """amax splits its gradient between tied channels; the fused backward tail takes maximum and tie count from the
    forward tail and must select the same channels."""
    r = G.case_convblock_ties(dtype)
    assert r["tie_share"] > 0.3, r["tie_share"]
    _check(r, dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [16, 32])
def test_head_mse_fused_matches_head_then_mse(G, C, dtype):
    """ops.HeadMSEFn (head 1x1 + nn.MSELoss in one pass each way) against ops.HeadFn followed by ops.mse_loss and against
    torch fp64 on the same (rounded) operands: loss, dx, dw, db."""
    from pcm_b200 import ops
    g = torch.Generator().manual_seed(5 + C)
    N, H, W, K = 5, 12, 18, 2
    x = torch.randn(N, H, W, C, generator=g).to(dtype)
    w = (torch.randn(K, C, 1, 1, generator=g) / C ** 0.5)
    b = torch.randn(K, generator=g)
    y = torch.randn(N, K, H, W, generator=g)
    res = {}
    for name in ("fused", "split", "torch"):
        if name == "torch":
            xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True); br = b.double().requires_grad_(True)
            pred = torch.einsum("nhwc,kc->nkhw", xr, wr[:, :, 0, 0]) + br[None, :, None, None]
            loss = ((pred - y.double()) ** 2).mean()
        else:
            xr = x.cuda().requires_grad_(True); wr = w.cuda().requires_grad_(True); br = b.cuda().requires_grad_(True)
            if name == "fused":
                assert ops.head_mse_ok(C, K)
                loss = ops.HeadMSEFn.apply(xr, wr, br, y.cuda())
            else:
                loss = ops.mse_loss(ops.HeadFn.apply(xr, wr, br), y.cuda())
        (loss * 3.0).backward()
        res[name] = [t.detach().double().cpu() for t in (loss, xr.grad, wr.grad, br.grad)]
    tol = 1e-5 if dtype == torch.float32 else 1e-2        # bf16: dx is rounded to bf16
    for a, f, t in zip(res["split"], res["fused"], res["torch"]):
        assert float((f - t).norm() / t.norm()) < tol
        assert float((f - a).norm() / t.norm()) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("tag", ["attunet_small", "attunet_cfg3_b2"])
def test_attunet(G, tag, dtype):
    _check(G.case_attunet(tag, dtype), dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simplecnn_small(G, dtype):
    """SimpleCNN (src/models.py:76-123), training-mode BatchNorm, against the oracle and the reference-made golden."""
    r = G.case_simplecnn(dtype)
    _check(r, dtype, median=SIMPLECNN_BF16_MEDIAN)
    assert r["running_mean"] < (1e-4 if dtype == torch.float32 else 2e-2), r["running_mean"]
    assert r["running_var"] < (1e-4 if dtype == torch.float32 else 2e-2), r["running_var"]
    assert r["nbt"] == 1


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cnn_transformer_small(G, dtype):
    """CNNTransformer (src/cnn_transformer.py), dropout 0, against the oracle and the reference-made golden."""
    _check(G.case_cnn_transformer(dtype), dtype)


@pytest.mark.parametrize("kind", ["simplecnn", "cnn_transformer"])
def test_full_size_tensor_core_paths(G, kind):
    """Default widths (init_dim 64 / embed_dim 128, head dim 32): 3x3, 1x1 and linear layers on tcgen05."""
    from pcm_b200._lib import lib
    r = G.case_full_size(kind)
    assert lib()._fn["pcm_tc_error_count"]() == 0
    _check(r, torch.bfloat16, median=SIMPLECNN_BF16_MEDIAN if kind == "simplecnn" else None)


def test_config5_geometry(G):
    """BASELINE configs[4]: 184x360 grid, base 32 — kernels beyond the shared-memory-resident regime."""
    from pcm_b200._lib import lib
    r = G.case_config5_geometry()
    assert lib()._fn["pcm_tc_error_count"]() == 0
    _check(r, torch.bfloat16)


def test_simplecnn_eval_mode(G):
    assert G.case_simplecnn_eval()["out"] < 1e-4


def test_dropout_masks(G):
    r = G.case_dropout_stats()
    assert abs(r["keep"] - 0.8) < 0.01 and r["same_mask"] and abs(r["scale"] - 1.25) < 1e-2, r
    assert r["whole_channels"] and abs(r["keep2d"] - 0.8) < 0.06, r


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_relu_dropout_fused_equals_three_steps(G, dtype):
    """linear1 -> ReLU -> dropout of the transformer FFN in one Function (bf16: in the GEMM's store epilogue,
    pcm_conv1x1_drop_tc): same mask as pcm_dropout with the same seed, same values up to one rounding, and a backward
    that needs only the saved output (pcm_relu_bwd_scaled)."""
    from pcm_b200 import ops_nn
    g = torch.Generator().manual_seed(3)
    B, L, K, N, p, seed = 3, 216, 128, 256, 0.3, 12345
    x = torch.randn(B, L, K, generator=g).to(dtype).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    dy = torch.randn(B, L, N, generator=g).to(dtype).cuda()
    outs = {}
    for name in ("fused", "steps"):
        xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        if name == "fused":
            y = ops_nn.LinearFn.apply(xr, wr, br, True, p, seed)
        else:
            y = ops_nn.DropoutFn.apply(ops_nn.LinearFn.apply(xr, wr, br, True), p, seed)
        y.backward(dy)
        outs[name] = [t.detach().float() for t in (y, xr.grad, wr.grad, br.grad)]
    yf, ys = outs["fused"][0], outs["steps"][0]
    assert torch.equal(yf == 0, ys == 0)                                   # same ReLU zeros and the same dropout mask
    keep = float(((ys != 0).sum() / (torch.relu(x.float() @ w.t() + b) > 0).sum()).item())
    assert abs(keep - (1 - p)) < 0.02, keep
    tol = 1e-5 if dtype == torch.float32 else 8e-3                         # bf16: one rounding instead of two
    for a, c in zip(outs["fused"], outs["steps"]):
        assert float((a - c).norm() / c.norm()) < tol


def test_metric_appendix_g(G):
    r = G.case_metric_appendix_g()            # fixture of _test_kaggle_metric.py:33-78, SURVEY Appendix G
    assert r["max_rel"] < 1e-5 and r["score_rel"] < 1e-5, r


def test_metric_full_size(G):
    r = G.case_metric(1080)                   # (1080, 2, 48, 72): the validation-set size of main_final.py
    assert r["max_rel"] < 1e-5 and r["score_rel"] < 1e-5, r


def test_metric_time_sharded(G):
    """Partial sums over time shards add (data-parallel validation): 3 ragged shards == one pass."""
    from oracle import metric_oracle as MO
    from pcm_b200 import metric as M
    pred, true, lat = MO.synth_metric_arrays(100)
    p, t = torch.from_numpy(pred).cuda(), torch.from_numpy(true).cuda()
    part = None
    for a, b in [(0, 37), (37, 38), (38, 100)]:
        part = M.metric_partial_sums(p[a:b], t[a:b], part)
    got = M.metric_finalize(part, lat, 100).cpu().numpy()
    w = MO.get_lat_weights(lat)
    for i in range(2):
        want = MO.metric_triplet(pred[:, i], true[:, i], w)
        np.testing.assert_allclose(got[i], want, rtol=1e-6)


def test_adam(G):
    assert G.case_adam()["max_rel"] < 1e-6


def test_season_stage(G):
    assert G.case_season_stage()["max_abs"] < 1e-6


def test_no_cpu_fallback(G):
    from pcm_b200 import ops
    from pcm_b200.src.unet import ConvBlock
    with pytest.raises(RuntimeError):
        ConvBlock(8, 16)(torch.zeros(1, 8, 8, 8))           # CPU tensor: must fail loudly
    with pytest.raises(RuntimeError):
        ops.mse_loss(torch.zeros(4), torch.zeros(4))


def test_train_step_graph_matches_eager_and_oracle(G):
    """3 optimisation steps through the captured CUDA graph == eager == oracle Adam steps (fp32 path)."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    B, T, H, W, base = 2, 3, 16, 24, 8
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 77)
    batches = [O.synth_attunet_batch(B, T, H, W, 78 + i)[:2] for i in range(3)]
    pcm_b200.set_compute_dtype(torch.float32)
    try:
        results = []
        for use_graph in (False, True):
            m = AttUNetConvLSTM(7, 2, base, seq_len=T)
            m.load_state_dict(sd)
            m = m.cuda()
            ts = TrainStep(m, (B, T, 7, H, W), (B, 2, H, W), lr=1e-3, use_graph=use_graph)
            ts.load_batch(batches[0][0].cuda(), batches[0][1].cuda())
            if use_graph:
                ts.warmup_and_capture(warmup=2)
                with torch.no_grad():                      # undo the warm-up updates
                    for k, p in m.named_parameters():
                        p.copy_(sd[k].cuda())
                ts.reset_optimizer_state()
            losses = [float(ts.step(x.cuda(), y.cuda()).item()) for x, y in batches]
            results.append((losses, {k: p.detach().cpu().clone() for k, p in m.named_parameters()}))
    finally:
        pcm_b200.set_compute_dtype(torch.bfloat16)
    # oracle: same three Adam steps on the CPU
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    used = {k: v for k, v in params.items() if not k.startswith("post_conv")}
    ms = {k: torch.zeros_like(v) for k, v in used.items()}
    vs = {k: torch.zeros_like(v) for k, v in used.items()}
    o_losses = []
    for it, (x, y) in enumerate(batches):
        for v in used.values():
            v.grad = None
        loss = O.mse_loss(O.attunet_convlstm(x, params), y)
        loss.backward()
        o_losses.append(float(loss))
        with torch.no_grad():
            for k, v in used.items():
                O.adam_step(v, v.grad, ms[k], vs[k], it + 1, lr=1e-3)
    for losses, ps in results:
        np.testing.assert_allclose(losses, o_losses, rtol=2e-4)
        bad, num, den = [], 0.0, 0.0
        for k, v in used.items():
            # Adam's first steps move every weight by ~lr*sign(g): an element whose gradient is ~0 can
            # flip sign between implementations, so compare the UPDATE in rel-L2, not element-wise.
            upd, want = ps[k] - sd[k], v.detach() - sd[k]
            num += float((upd - want).norm()) ** 2
            den += float(want.norm()) ** 2
            e = float((upd - want).norm() / want.norm())
            bad.append((k, e)) if e > 0.5 else None
        assert not bad, bad
        assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5      # all parameters together
        assert torch.equal(ps["post_conv.0.weight"], sd["post_conv.0.weight"])     # untouched (zero grad, F5)
    np.testing.assert_allclose(results[0][0], results[1][0], rtol=1e-5)


def test_step_prefetch_matches_step(G):
    """TrainStep.step_prefetch (next batch copied host->device on a second stream while the current step runs) yields
    the same loss sequence as the serial step() on the same batches, under the captured graph."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    B, T, H, W, base = 2, 3, 16, 24, 8
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 81)
    batches = [tuple(t.pin_memory() for t in O.synth_attunet_batch(B, T, H, W, 82 + i)[:2]) for i in range(4)]
    pcm_b200.set_compute_dtype(torch.float32)
    try:
        losses = []
        for prefetch in (False, True):
            model = AttUNetConvLSTM(7, 2, base, seq_len=T)
            model.load_state_dict(sd)
            model = model.cuda()
            step = TrainStep(model, (B, T, 7, H, W), (B, 2, H, W), lr=1e-3)
            step.load_batch(*batches[0])
            step.warmup_and_capture(warmup=2)
            with torch.no_grad():                         # undo the warm-up updates
                for k, p in model.named_parameters():
                    p.copy_(sd[k].cuda())
            step.reset_optimizer_state()
            out = []
            if prefetch:
                step.load_batch(*batches[0])
                for i in range(4):
                    out.append(float(step.step_prefetch(*batches[(i + 1) % 4]).item()))
            else:
                for i in range(4):
                    out.append(float(step.step(*batches[i]).item()))
            losses.append(out)
    finally:
        pcm_b200.set_compute_dtype(torch.bfloat16)
    for a, b in zip(*losses):
        assert abs(a - b) / abs(a) < 1e-4, losses


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_side_stream_matches_single_stream(G, dtype, monkeypatch):
    """The weight-gradient kernels run on a second stream inside the captured step (fork / join around every ConvBlock,
    ConvTranspose and ConvLSTM backward): losses and final parameters must equal the single-stream step's (same
    kernels, same inputs; only fp32 atomics may reorder)."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    B, T, H, W, base = 4, 3, 48, 72, 16            # full-size grid: the per-image tails and the tensor-core kernels
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 91)
    batches = [O.synth_attunet_batch(B, T, H, W, 92 + i)[:2] for i in range(3)]
    pcm_b200.set_compute_dtype(dtype)
    try:
        runs = []
        for side in ("0", "1"):
            monkeypatch.setenv("PCM_SIDE_STREAM", side)
            model = AttUNetConvLSTM(7, 2, base, seq_len=T)
            model.load_state_dict(sd)
            model = model.cuda()
            step = TrainStep(model, (B, T, 7, H, W), (B, 2, H, W), lr=1e-3)
            assert (step.side is not None) == (side == "1")
            step.load_batch(*batches[0])
            step.warmup_and_capture(warmup=2)
            with torch.no_grad():                         # undo the warm-up updates
                for k, p in model.named_parameters():
                    p.copy_(sd[k].cuda())
            step.reset_optimizer_state()
            losses = [float(step.step(*batches[i % 3]).item()) for i in range(6)]
            flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
            runs.append((losses, flat))
    finally:
        pcm_b200.set_compute_dtype(torch.bfloat16)
    (l0, p0), (l1, p1) = runs
    tol = 1e-4 if dtype == torch.float32 else 5e-3
    for a, b in zip(l0, l1):
        assert abs(a - b) / abs(a) < tol, (l0, l1)
    assert float((p0 - p1).norm() / p0.norm()) < tol


@pytest.mark.parametrize("switch", ["PCM_BUCKETS_SINGLE", "PCM_HEAD_MSE", "PCM_CONV_GROUP", "PCM_WGRAD_GROUP", "PCM_TAIL_WAVES",
                                    "PCM_POOL_SDOT", "PCM_TAIL_SCRATCH"])
def test_step_variants_agree(G, switch, monkeypatch):
    """Every scheduling / kernel-form switch of the captured step — gradient buckets with per-bucket fold + Adam on the
    communication stream at world size 1, the fused head + loss, the pixel-group forms of the thin convolutions and weight
    gradients, the per-launch CTA width of the tails, the gate-gradient sum handed from the pooling backward to the block
    backward, the shared-memory scratch of the level-1 backward tails — must leave losses and parameters where the plain variant puts them
    (same mathematics; only fp32 summation order and bf16 rounding of equal sums may differ)."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    B, T, H, W, base = 4, 3, 48, 72, 16
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 191)
    batches = [O.synth_attunet_batch(B, T, H, W, 192 + i)[:2] for i in range(3)]
    runs = []
    for val in ("0", None):
        if val is None:
            monkeypatch.delenv(switch, raising=False)
        else:
            monkeypatch.setenv(switch, val)
        model = AttUNetConvLSTM(7, 2, base, seq_len=T)
        model.load_state_dict(sd)
        model = model.cuda()
        step = TrainStep(model, (B, T, 7, H, W), (B, 2, H, W), lr=1e-3)
        if switch == "PCM_BUCKETS_SINGLE":
            assert bool(step.buckets) == (val is None)
        step.load_batch(*batches[0])
        step.warmup_and_capture(warmup=2)                  # parameters / Adam state restored afterwards
        losses = [float(step.step(*batches[i % 3]).item()) for i in range(6)]
        step.check_kernels()
        flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
        runs.append((losses, flat))
    (l0, p0), (l1, p1) = runs
    for a, b in zip(l0, l1):
        assert abs(a - b) / abs(a) < 5e-3, (l0, l1)
    assert float((p0 - p1).norm() / p0.norm()) < 5e-3


@pytest.mark.parametrize("kind", ["simplecnn", "cnn_transformer", "unet"])
def test_trainstep_captures_every_model_family(G, kind):
    """TrainStep (CUDA graph, second stream enabled by default) on the models that issue little or no side-stream work:
    the capture must stay valid (joining a stream that was never forked would invalidate it) and the loss must fall."""
    import pcm_b200
    from pcm_b200.trainer import TrainStep
    torch.manual_seed(0)
    if kind == "simplecnn":
        from pcm_b200.src.models import SimpleCNN
        model, H, W = SimpleCNN(5, 2, kernel_size=3, init_dim=16, depth=2, dropout_rate=0.0), 16, 24
    elif kind == "cnn_transformer":
        from pcm_b200.src.cnn_transformer import CNNTransformer
        model, H, W = CNNTransformer(5, 2, 32, 2, 4, 64, dropout=0.0), 48, 72
    else:
        from pcm_b200.src.unet import UNet
        model, H, W = UNet(5, 2, 16), 16, 24
    model = model.cuda()
    B = 4
    x, y = torch.randn(B, 5, H, W, device="cuda"), torch.randn(B, 2, H, W, device="cuda")
    step = TrainStep(model, (B, 5, H, W), (B, 2, H, W), lr=2e-3)
    assert step.side is not None
    step.load_batch(x, y)
    step.warmup_and_capture(warmup=2)
    assert step.graph is not None
    losses = [float(step.step(x, y).item()) for _ in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses


def test_forward_windows_matches_stacked_windows(G):
    """Device-resident window gather (SequenceDataset semantics, zero left-pad) == forward on the explicitly stacked
    windows; includes target indices smaller than seq_len - 1."""
    import pcm_b200
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    T, H, W, base, Ttot = 4, 16, 24, 8, 12
    sd = O.synth_state_dict(O.attunet_spec(7, 2, base), 91)
    g = torch.Generator().manual_seed(92)
    series = torch.randn(Ttot, 7, H, W, generator=g)
    idx = torch.tensor([0, 2, 3, 7, 11])
    pad = torch.zeros(7, H, W)
    x_seq = torch.stack([torch.stack([series[i - T + 1 + t] if i - T + 1 + t >= 0 else pad for t in range(T)]) for i in idx.tolist()])
    pcm_b200.set_compute_dtype(torch.float32)
    try:
        model = AttUNetConvLSTM(7, 2, base, seq_len=T)
        model.load_state_dict(sd)
        model = model.cuda()
        with torch.no_grad():
            a = model.forward_windows(series.cuda(), idx.cuda()).cpu()
            b = model(x_seq.cuda()).cpu()
    finally:
        pcm_b200.set_compute_dtype(torch.bfloat16)
    assert float((a - b).abs().max()) < 1e-5            # same kernels; only the order of fp32 atomic sums differs
    ref = O.attunet_convlstm(x_seq.double(), {k: v.double() for k, v in sd.items()})
    assert float((a.double() - ref).norm() / ref.norm()) < 1e-4


def test_metric_with_fused_inverse_transform(G):
    """Normalised pred/truth + Normalizer.inverse_transform_output fused into the accumulation (zscore for tas, log1p for
    pr — the reference's default output transforms) == the metric oracle on the de-normalised fp64 arrays."""
    from oracle import metric_oracle as MO
    from pcm_b200 import metric as M
    pred, true, lat = MO.synth_metric_arrays(120)                       # physical units, (T, 2, Y, X)
    stats = {0: {"method": "zscore", "params": {"mean": 280.0, "std": 12.0}},
             1: {"method": "log1p", "params": {"mean": 0.9, "std": 0.6}}}

    def norm(a):
        out = np.empty_like(a, dtype=np.float64)
        out[:, 0] = (a[:, 0].astype(np.float64) - 280.0) / 12.0
        out[:, 1] = (np.log1p(np.maximum(a[:, 1].astype(np.float64), 0.0)) - 0.9) / 0.6
        return out.astype(np.float32)

    pn, tn = norm(pred), norm(true)
    # what the reference de-normalises to (fp64 on the fp32 normalised values)
    def denorm(a):
        out = np.empty(a.shape, dtype=np.float64)
        out[:, 0] = a[:, 0].astype(np.float64) * 12.0 + 280.0
        out[:, 1] = np.expm1(a[:, 1].astype(np.float64) * 0.6 + 0.9)
        return out
    w = MO.get_lat_weights(lat)
    dp, dt = denorm(pn), denorm(tn)
    table = M.transform_table(stats, 2, "cuda")
    part = M.metric_partial_sums_normalized(torch.from_numpy(pn).cuda(), torch.from_numpy(tn).cuda(), table)
    got = M.metric_finalize(part, lat, 120).cpu().numpy()
    for i in range(2):
        want = MO.metric_triplet(dp[:, i], dt[:, i], w)
        np.testing.assert_allclose(got[i], want, rtol=1e-5)


def test_metric_skips_nans_like_xarray(G):
    """NaN cells (e.g. the tas < 150 K placeholders the reference masks, main_final.py:222-224) drop out of the time
    mean / std of their pixel and of the weighted means (src/utils_final.py:296) — against the numpy restatement."""
    from oracle import metric_oracle as MO
    from pcm_b200 import metric as M
    pred, true, lat = MO.synth_metric_arrays(60)
    rs = np.random.RandomState(0)
    pred, true = pred.copy(), true.copy()
    true[rs.rand(*true.shape) < 0.02] = np.nan            # scattered missing targets
    pred[rs.rand(*pred.shape) < 0.01] = np.nan
    true[:, 0, 5, 7] = np.nan                              # a pixel with no valid target at all
    w = MO.get_lat_weights(lat)
    got = M.weighted_metric_triplets(torch.from_numpy(pred).cuda(), torch.from_numpy(true).cuda(), lat)
    for i in range(2):
        np.testing.assert_allclose(got[i], MO.metric_triplet(pred[:, i], true[:, i], w), rtol=1e-9)


def test_kaggle_score_matches_unmodified_reference(G, golden_dir):
    """a18: the Kaggle form of the score (weights cos(lat_2dp)/sum, _climate_kaggle_metric.py:103-153) on the device,
    against numbers the UNMODIFIED reference scorer produced (oracle/make_goldens.py): the Appendix G fixture through the
    array route, and a DataFrame round trip (our writer -> our parser -> device reductions)."""
    import json
    import os
    from oracle import kaggle_oracle as KO
    from oracle import metric_oracle as MO
    from pcm_b200 import kaggle as K
    with open(os.path.join(golden_dir, "metric_appendix_g.json")) as f:
        want = json.load(f)["reference_kaggle_score"]
    fx = MO.known_answer_fixture()
    got = K.score_arrays({v: fx[v + "_pred"] for v in ("tas", "pr")}, {v: fx[v + "_true"] for v in ("tas", "pr")}, fx["lats"])
    assert abs(got - want) / want < 1e-9, (got, want)
    with open(os.path.join(golden_dir, "kaggle_roundtrip.json")) as f:
        g = json.load(f)
    pred, true, lat, lon, names = KO.synth_submission(T=g["T"], seed=g["seed"])
    sol32 = K.convert_predictions_to_kaggle_format(true, np.arange(g["T"]), lat, lon, names)
    sub32 = K.convert_predictions_to_kaggle_format(pred, np.arange(g["T"]), lat, lon, names)
    # float32 columns (in memory): the reference's numpy reductions then run in float32 — agreement to float32 rounding
    got = K.score(sol32, sub32, "ID")
    assert abs(got - g["reference_score_float32_columns"]) / got < 2e-6, (got, g["reference_score_float32_columns"])
    # float64 columns (what a scorer reads back from the submission CSV): agreement to fp64 rounding
    sol, sub = sol32.astype({"Prediction": np.float64}), sub32.astype({"Prediction": np.float64})
    got = K.score(sol, sub, "ID")
    assert abs(got - g["reference_score"]) / g["reference_score"] < 1e-9, (got, g["reference_score"])
    perm = np.random.RandomState(3).permutation(len(sub))
    got = K.score(sol, sub.iloc[perm].reset_index(drop=True), "ID")
    assert abs(got - g["reference_score_shuffled_submission"]) / g["reference_score"] < 1e-9
    with pytest.raises(ValueError, match="Submission must have columns"):
        K.score(sol, sub.rename(columns={"Prediction": "p"}), "ID")
    with pytest.raises(ValueError, match="missing predictions"):
        K.score(sol, sub.iloc[:-5], "ID")


def test_window_stage_fused_normalizer_and_season(G):
    """(f)2: raw record in HBM -> Normalizer.normalize (zscore / minimax / log1p / sqrt / pow / pass-through) + seasonal
    sin/cos channels + zero left-pad in the staging kernel == the reference's data pipeline order (normalise the record,
    append the month channels, slice windows, pad with zeros)."""
    import pcm_b200
    from pcm_b200 import ops
    from pcm_b200.data import Normalizer
    Ttot, C, H, W, T = 14, 5, 8, 12, 4
    g = torch.Generator().manual_seed(5)
    raw = torch.rand(Ttot, C, H, W, generator=g) * 3 + 0.1
    stats = {0: {"method": "zscore", "params": {"mean": 1.5, "std": 0.8}},
             1: {"method": "minimax", "params": {"min_val": 0.1, "max_val": 3.1}},
             2: {"method": "log1p", "params": {"mean": 0.9, "std": 0.4}},
             3: {"method": "sqrt", "params": {"mean": 1.2, "std": 0.3}},
             4: {"method": "pow", "params": {"lambda": 0.25, "mean": 1.1, "std": 0.2}}}
    nz = Normalizer()
    nz.set_input_statistics(stats)
    month = torch.arange(Ttot, dtype=torch.int32) % 12
    # reference order of operations, numpy fp64 (src/utils_final.py:76-128, main_final.py:188-216)
    r = raw.double().numpy()
    n = np.stack([(r[:, 0] - 1.5) / (0.8 + 1e-8), (r[:, 1] - 0.1) / 3.0, (np.log1p(r[:, 2]) - 0.9) / (0.4 + 1e-8),
                  (np.sqrt(r[:, 3]) - 1.2) / (0.3 + 1e-8), (r[:, 4] ** 0.25 - 1.1) / (0.2 + 1e-8)], 1)
    ang = 2 * np.pi * month.numpy() / 12
    rec = np.concatenate([n, np.broadcast_to(np.sin(ang)[:, None, None, None], (Ttot, 1, H, W)),
                          np.broadcast_to(np.cos(ang)[:, None, None, None], (Ttot, 1, H, W))], 1).astype(np.float32)
    idx = torch.tensor([0, 2, 3, 9, 13])
    want = np.zeros((T, len(idx), H, W, 16), np.float32)
    for b, i in enumerate(idx.tolist()):
        for t in range(T):
            f = i - T + 1 + t
            if f >= 0:
                want[t, b, :, :, :7] = rec[f].transpose(1, 2, 0)
    got = ops.window_stage(raw.cuda(), idx.cuda(), T, torch.float32, norm=nz.input_table(C, "cuda"),
                           month=month.cuda()).cpu().numpy().reshape(T, len(idx), H, W, 16)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    # stand-alone Normalizer.normalize on the device
    out = nz.normalize(raw.cuda(), "input").cpu().numpy()
    np.testing.assert_allclose(out, n.astype(np.float32), rtol=2e-6, atol=2e-6)
