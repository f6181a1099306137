"""Micro-benchmark of the ConvLSTM recurrence at the benchmark shape (T=6, B=64, 128 -> 64 channels, 6 x 9), forward +
backward under a CUDA graph: persistent cluster kernels (csrc/convlstm_seq.cu) against the per-step launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcm_b200  # noqa: E402,F401
from pcm_b200 import ops  # noqa: E402
from pcm_b200._lib import lib  # noqa: E402

T, B, Cin, Ch, H, W = 6, 64, 128, 64, 6, 9


def bench(persistent, only=None):
    os.environ["PCM_LSTM_PERSISTENT"] = "1" if persistent else "0"
    torch.manual_seed(0)
    w = (torch.randn(4 * Ch, Cin + Ch, 3, 3, device="cuda") * 0.02).requires_grad_(True)
    b = torch.zeros(4 * Ch, device="cuda", requires_grad=True)
    x = torch.randn(T * B, H, W, Cin, device="cuda").bfloat16().requires_grad_(True)
    gy = torch.randn(B, H, W, Ch, device="cuda").bfloat16()
    plan = ops.PackPlan()

    def step():
        with ops.use_pack_plan(plan):
            plan.repack()
            h = ops.ConvLSTMFn.apply(x, w, b, T, B, B, 1, True)
            if only != "fwd":
                h.backward(gy)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            n0 = lib().launches
            step()
            n = lib().launches - n0
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 50 * 1e3, n


for only in ("fwd", None):
    for persistent in (False, True):
        us, n = bench(persistent, only)
        print(f"{'fwd only' if only else 'fwd+bwd '}  persistent={int(persistent)}  {us:8.1f} us per replay  ({n} launches)")
print("tc errors", lib()._fn["pcm_tc_error_count"]())
