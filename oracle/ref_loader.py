"""By-path loader for the REAL reference modules (TEST INFRASTRUCTURE; build container only).

`/root/reference` exists only in the build container, never on the GPU box, so this module is
used solely by ``oracle/make_goldens.py`` and by CPU tests that skip when the tree is absent.
It reads the reference in place (nothing is copied into this repo) under a synthetic package
name so the relative imports at src/unet_convlstm_attention.py:10-11 resolve without executing
src/__init__.py (which needs omegaconf) — SURVEY.md Appendix F.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("PCM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "convlstm.py"))


def load():
    """Returns a namespace with the reference's convlstm / unet / unet_convlstm_attention /
    cnn_transformer / models modules and the unmodified kaggle ``score``."""
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    src = os.path.join(REF_ROOT, "src")
    if "refsrc" not in sys.modules:
        pkg = types.ModuleType("refsrc")
        pkg.__path__ = [src]
        sys.modules["refsrc"] = pkg
        if "omegaconf" not in sys.modules:           # only a type hint at src/models.py:2,7
            om = types.ModuleType("omegaconf")
            om.DictConfig = dict
            sys.modules["omegaconf"] = om
        for name in ["convlstm", "unet", "unet_convlstm_attention", "cnn_transformer", "models"]:
            spec = importlib.util.spec_from_file_location(f"refsrc.{name}", os.path.join(src, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[f"refsrc.{name}"] = mod
            spec.loader.exec_module(mod)
    ns = types.SimpleNamespace(**{n: sys.modules[f"refsrc.{n}"] for n in
                                  ["convlstm", "unet", "unet_convlstm_attention", "cnn_transformer", "models"]})
    spec = importlib.util.spec_from_file_location("ref_kaggle_metric", os.path.join(REF_ROOT, "_climate_kaggle_metric.py"))
    km = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(km)
    ns.score = km.score
    return ns
