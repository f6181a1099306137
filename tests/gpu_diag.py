"""Run every GPU parity case and print a table (does not stop at failures).  Usage on the GPU box:
    python tests/gpu_diag.py [> gpurun_out/diag.txt]"""
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_cases as G  # noqa: E402


def show(name, fn, *a):
    try:
        r = fn(*a)
    except Exception:
        print(f"[{name}] EXCEPTION\n{traceback.format_exc()}")
        return
    flat = {k: v for k, v in r.items() if not isinstance(v, dict)}
    print(f"[{name}] " + " ".join(f"{k}={v:.3e}" for k, v in flat.items()))
    gn = r.get("gnorm", {})
    for key in ("grads", "golden_grads"):
        if key in r:
            worst = sorted(r[key].items(), key=lambda kv: -kv[1])[:6]
            print(f"    worst {key}: " + ", ".join(f"{k}={v:.2e}(|g|={gn.get(k, float('nan')):.1e})" for k, v in worst))
    sys.stdout.flush()


def main():
    print(torch.cuda.get_device_name(0))
    for dt in (torch.float32, torch.bfloat16):
        tag = "f32" if dt == torch.float32 else "bf16"
        show(f"se/{tag}", G.case_se, dt)
        show(f"spatial_gate/{tag}", G.case_spatial_gate, dt)
        show(f"convblock/{tag}", G.case_convblock, dt)
        show(f"down_up/{tag}", G.case_down_up, dt)
        show(f"cell_step/{tag}", G.case_cell_step, dt)
        show(f"convlstm/{tag}", G.case_convlstm, dt)
        show(f"unet/{tag}", G.case_unet, dt)
        show(f"attunet_small/{tag}", G.case_attunet, "attunet_small", dt)
        show(f"attunet_cfg3_b2/{tag}", G.case_attunet, "attunet_cfg3_b2", dt)
    show("metric_appendix_g", G.case_metric_appendix_g)
    show("metric_1080", G.case_metric, 1080)
    show("adam", G.case_adam)
    show("season_stage", G.case_season_stage)


if __name__ == "__main__":
    main()
