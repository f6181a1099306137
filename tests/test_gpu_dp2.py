"""GPU (needs 2 devices; skipped otherwise): the data-parallel training step with the overlapped all-reduce buckets
(three by default: [ConvLSTM, decoder, head] / [enc2..enc4] / [enc1], each followed by its own Adam range update).  Rank r trains on half r of a batch; samples are independent (GroupNorm is per sample) and the loss is a
mean, so (sum of the two ranks' gradients) / 2 must equal the gradient a single process computes on the whole batch
— checked on the flat gradient buffer after one step (a gradient left out of a bucket would not be summed), on the
eager path and under the captured graph."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(rank, world, port, use_graph, fp32, out, lr=0.0, nsteps=1):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import pcm_b200  # noqa: F401
    from oracle import model_oracle as O
    from pcm_b200.src.unet_convlstm_attention import AttUNetConvLSTM
    from pcm_b200.trainer import TrainStep
    from pcm_b200.config import set_compute_dtype
    if fp32:
        set_compute_dtype(torch.float32)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    Bt, T, H, W = 8, 3, 16, 24
    B = Bt // world
    sd = O.synth_state_dict(O.attunet_spec(7, 2, 16), 21)
    model = AttUNetConvLSTM(7, 2, 16, seq_len=T)
    model.load_state_dict(sd)
    model = model.to(dev)
    step = TrainStep(model, (B, T, 7, H, W), (B, 2, H, W), lr=lr, use_graph=use_graph)   # lr 0: weights stay put
    assert world == 1 or step.split > 0
    x, y, _ = O.synth_attunet_batch(Bt, T, H, W, 22)
    step.load_batch(x[rank * B:(rank + 1) * B].to(dev), y[rank * B:(rank + 1) * B].to(dev))
    step.warmup_and_capture(warmup=2)              # restores parameters / Adam state afterwards
    for _ in range(nsteps):
        step.run()
    torch.cuda.synchronize()
    if rank == 0:
        ranges = [(lo, hi) for _, lo, hi in step.buckets] + ([step.tail_bucket] if step.tail_bucket else [])
        torch.save({"grad": (step.opt.flat_grad / world).cpu(), "split": step.split, "n": step.opt.n_reduced,
                    "ranges": ranges, "param": step.opt.flat_param.cpu(), "adam_step": float(step.opt.state[0])}, out)
    # leave without any further collective and without tearing the communicator down: destroying a NCCL group whose
    # kernels a captured graph still references can hang (same exit path as bench.py)
    step.graph = None
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


@pytest.mark.timeout(240)
@pytest.mark.parametrize("use_graph,fp32", [(False, True), (True, False)])
def test_two_bucket_allreduce_matches_single_process(tmp_path, use_graph, fp32):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ref, dp = str(tmp_path / "ref.pt"), str(tmp_path / "dp.pt")
    mp.spawn(_run, args=(1, 0, use_graph, fp32, ref), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), use_graph, fp32, dp), nprocs=2, join=True)
    a, b = torch.load(ref), torch.load(dp)
    split, n = b["split"], b["n"]
    assert 0 < split < n
    ranges = sorted(b["ranges"])
    assert len(ranges) == 3 and ranges[0][0] == 0 and ranges[-1][1] == n                 # the buckets tile [0, n_reduced)
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(2))
    errs = []
    for lo, hi in ranges:
        ga, gb = a["grad"][lo:hi], b["grad"][lo:hi]
        errs.append(float((ga - gb).norm() / ga.norm()))
    print("bucket errors", errs)
    # fp32 storage: only the order of fp32 atomic accumulations differs -> tight.  bf16 storage: last-bit differences
    # in the (atomically accumulated) GroupNorm statistics flip bf16 roundings downstream, so two runs of the SAME
    # configuration already differ by ~1e-2 in the gradients; a bucket that missed its sum would be off by > 0.3.
    assert max(errs) < (1e-4 if fp32 else 0.1), errs


@pytest.mark.timeout(240)
def test_per_bucket_adam_matches_single_process(tmp_path):
    """lr > 0: after three steps on the same (half) batches the parameters of the data-parallel run — whose Adam updates
    run bucket by bucket on the communication stream — equal the single-process run's; the step counter advanced once
    per step."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ref, dp = str(tmp_path / "ref.pt"), str(tmp_path / "dp.pt")
    mp.spawn(_run, args=(1, 0, True, True, ref, 1e-3, 3), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), True, True, dp, 1e-3, 3), nprocs=2, join=True)
    a, b = torch.load(ref), torch.load(dp)
    assert a["adam_step"] == 3.0 and b["adam_step"] == 3.0
    n = b["n"]
    err = float((a["param"][:n] - b["param"][:n]).norm() / a["param"][:n].norm())
    print("parameter error after 3 steps", err)
    assert err < 1e-4, err
