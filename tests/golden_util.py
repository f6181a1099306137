"""Shared helpers: load a golden fixture and compare a {name: grad} dict against it."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 1024


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    return cfg, z


def rel_l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def grad_errors(named_grads, z, floor=1e-7):
    """rel-L2 of each parameter gradient's strided sample (and its L2 norm) against the golden.
    Returns {name: (sample_rel_l2, norm_rel, golden_norm)}; gradients the reference leaves None
    must be None/zero here."""
    out = {}
    for key in z.files:
        if key.startswith("gradnone/"):
            k = key[len("gradnone/"):]
            g = named_grads.get(k)
            assert g is None or float(torch.as_tensor(g).abs().max()) == 0.0, f"{k}: reference grad is None"
        if not key.startswith("grad/"):
            continue
        k = key[len("grad/"):]
        g = torch.as_tensor(named_grads[k]).detach().float().cpu().reshape(-1)
        stride = max(1, g.numel() // SAMPLE)
        ref = z[key]
        gn = float(z["gradnorm/" + k][0])
        if gn < floor:                      # dead unit (e.g. SE bottleneck relu): absolute check only
            out[k] = (float(g.norm()), 0.0, gn)
            continue
        out[k] = (rel_l2(g[::stride].numpy(), ref), abs(float(g.norm()) - gn) / gn, gn)
    return out
