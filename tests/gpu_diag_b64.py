"""Per-tensor error table of the benchmark-shape cases (calibration of the gates; run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_cases as G  # noqa: E402

for tag in ("attunet_cfg3_b64", "attunet_cfg3_b64_default_init"):
    for dt in (torch.float32, torch.bfloat16):
        r = G.case_attunet_b64(tag, dt)
        errs = {k: e for k, e in r["grads"].items() if r["gnorm"][k] > 1e-7}
        print(f"[{tag}/{'f32' if dt == torch.float32 else 'bf16'}] out={r['out']:.3e} out_norm={r['out_norm']:.3e} "
              f"loss={r['loss']:.3e} grad median={np.median(list(errs.values())):.3e} max={max(errs.values()):.3e}")
        for k, e in sorted(errs.items(), key=lambda kv: -kv[1])[:14]:
            print(f"      {e:.3e}  |g|={r['gnorm'][k]:.2e}  {k}")
        sys.stdout.flush()
