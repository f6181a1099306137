"""Train-step time of the other model variants that share the conv stack (BASELINE configs[0], configs[1], plus the
single-frame UNet) — not the headline benchmark; eager and CUDA-graph timings with CUDA events.

    python tools/bench_models.py [--steps 10]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcm_b200  # noqa: E402,F401
from pcm_b200.src.cnn_transformer import CNNTransformer  # noqa: E402
from pcm_b200.src.models import SimpleCNN  # noqa: E402
from pcm_b200.src.unet import UNet  # noqa: E402
from pcm_b200.trainer import TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--trace", action="store_true", help="per-kernel totals of one eager step (CUDA events per C-ABI call)")
    a = ap.parse_args()
    torch.manual_seed(42)
    cases = [
        ("SimpleCNN (init_dim 64, depth 4, dropout 0.2), batch 32", lambda: SimpleCNN(5, 2), 32, 222.239),
        ("cnn_transformer (embed 128, depth 4, 4 heads, dropout 0.1), batch 64", lambda: CNNTransformer(), 64, 1.157898),
        ("unet (base 16), batch 64", lambda: UNet(5, 2, 16), 64, None),
    ]
    for name, mk, B, gflop in cases:
        model = mk().cuda().train()
        step = TrainStep(model, (B, 5, 48, 72), (B, 2, 48, 72), lr=5e-4)
        x = torch.randn(B, 5, 48, 72, device="cuda")
        y = torch.randn(B, 2, 48, 72, device="cuda")
        step.load_batch(x, y)
        step.warmup_and_capture(warmup=3)
        for _ in range(3):
            step.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step.run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        extra = f"  {B / ms * gflop:8.1f} TFLOP/s algorithmic" if gflop else ""
        print(f"{name}: {ms:8.3f} ms/step  {B / ms * 1e3:9.0f} samples/s  {step.launches_per_step} launches{extra}  "
              f"loss {float(step.loss):.4f}")
        if a.trace:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from bench import kernel_trace
            _, agg = kernel_trace(step._step_impl, n_steps=2)
            for k, (ms_k, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
                print(f"    {ms_k * 1e3:9.1f} us  x{n:5.1f}  {k}")
        del step, model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
