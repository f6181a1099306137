"""GPU-resident input pipeline: the reference's Normalizer statistics
(src/utils_final.py:32-206) as device tables, and a window loader that keeps the whole input record in HBM
(8 109 x 5 x 48 x 72 fp32 = 560 MB) so that a training step needs B window indices from the host instead of a
(B, T, C, H, W) batch — sliding-window gather with zero left-pad (main_final.py:97-154), per-variable normalisation
(src/utils_final.py:45-128) and the sin/cos month channels (main_final.py:186-216) are fused into ONE staging kernel
(pcm_window_stage) that writes the first convolution's NHWC bf16 operand."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import metric, ops
from ._lib import lib

_FWD_KINDS = {"zscore": 0, "minimax": 0, "log1p": 1, "sqrt": 2, "pow": 3}
EPSILON = 1e-8                                   # src/utils_final.py:60


class Normalizer:
    """Same statistics contract as the reference's Normalizer: index-keyed maps
    {var_idx: {"method": "zscore"|"minimax"|"log1p"|"sqrt"|"pow", "params": {...}}}; variables without an entry pass
    through.  `normalize` / `inverse_transform_output` take CUDA tensors (N, C, H, W) and run on the device; the
    training path never calls them — it uses `input_table()` fused into the window staging and `output_table()` fused
    into the metric accumulation."""

    def __init__(self):
        self.input_stats = {}
        self.output_stats = {}

    def set_input_statistics(self, transform_map_indexed):
        self.input_stats = transform_map_indexed

    def set_output_statistics(self, transform_map_indexed):
        self.output_stats = transform_map_indexed

    # ---- device tables -------------------------------------------------------------------------------------
    @staticmethod
    def _forward_rows(stats: dict, n: int):
        rows = []
        for v in range(n):
            cfg = stats.get(v)
            if cfg is None:
                rows.append([-1.0, 1.0, 0.0, 1.0])
                continue
            m, pr = cfg["method"], cfg.get("params", {})
            if m not in _FWD_KINDS:
                raise ValueError(f"Unknown method '{m}' for var {v}.")
            if m == "minimax":
                if pr.get("min_val") is None or pr.get("max_val") is None:
                    raise ValueError(f"Minimax params missing for var {v}.")
                rng = float(pr["max_val"]) - float(pr["min_val"])
                scale = rng if not np.isclose(rng, 0) else 1.0
                rows.append([0.0, 1.0 / scale, float(pr["min_val"]), 1.0])
                continue
            if pr.get("mean") is None or pr.get("std") is None or (m == "pow" and pr.get("lambda") is None):
                raise ValueError(f"{m} params missing for var {v}.")
            rows.append([float(_FWD_KINDS[m]), 1.0 / (float(pr["std"]) + EPSILON), float(pr["mean"]),
                         float(pr.get("lambda", 1.0))])
        return rows

    def input_table(self, n_channels: int, device) -> torch.Tensor:
        """(C, 4) fp64 = (kind, a, b, lambda) for pcm_window_stage: x_n = (g(x) - b) * a."""
        if not self.input_stats:
            raise RuntimeError("Statistics for 'input' not set.")
        return torch.tensor(self._forward_rows(self.input_stats, n_channels), dtype=torch.float64, device=device)

    def output_table(self, n_vars: int, device) -> torch.Tensor:
        """(V, 4) fp32 for pcm_metric_partial_denorm (inverse transform fused into the metric accumulation)."""
        if not self.output_stats:
            raise RuntimeError("Output stats not set.")
        return metric.transform_table(self.output_stats, n_vars, device)

    # ---- stand-alone transforms on the device -----------------------------------------------------------------
    def normalize(self, data: torch.Tensor, data_type: str = "input") -> torch.Tensor:
        """(N, C, H, W) fp32 CUDA -> normalised fp32 (src/utils_final.py:45-128), via the staging kernel."""
        stats = self.input_stats if data_type == "input" else self.output_stats
        if not stats:
            raise RuntimeError(f"Statistics for '{data_type}' not set.")
        if not (isinstance(data, torch.Tensor) and data.is_cuda):
            raise TypeError("pcm_b200 Normalizer works on CUDA tensors (no CPU fallback)")
        N, C, H, W = data.shape
        table = torch.tensor(self._forward_rows(stats, C), dtype=torch.float64, device=data.device)
        idx = torch.arange(N, device=data.device)
        y = ops.window_stage(data, idx, 1, torch.float32, norm=table)                        # (N, H, W, Cp) fp32
        out = torch.empty((N, C, H, W), device=data.device, dtype=torch.float32)
        lib().call("pcm_nhwc_to_nchw", y.data_ptr(), out.data_ptr(), N, C, H, W, y.shape[-1], 1, 0,
                   torch.cuda.current_stream().cuda_stream)
        return out


class WindowLoader:
    """The training inputs resident in HBM.  `series` (Ttot, C, H, W) raw (or already normalised) fp32 forcings;
    `targets` (Ttot, V, H, W) normalised targets; optional `month` (Ttot,) month index 0..11 for the seasonal channels;
    optional Normalizer for the inputs.  `stage(idx)` -> (frames NHWC t-major for AttUNetConvLSTM.forward_staged,
    targets (B, V, H, W))."""

    def __init__(self, series: torch.Tensor, targets: torch.Tensor, seq_len: int, normalizer: Optional[Normalizer] = None,
                 month: Optional[torch.Tensor] = None):
        if not series.is_cuda:
            raise RuntimeError("WindowLoader keeps the record on the GPU (no CPU fallback)")
        self.series = series.contiguous().float()
        self.targets = targets.contiguous().float()
        self.seq_len = seq_len
        self.norm = normalizer.input_table(series.shape[1], series.device) if normalizer is not None else None
        self.month = month.to(device=series.device, dtype=torch.int32).contiguous() if month is not None else None

    @property
    def h2d_bytes_per_step(self):
        return 8          # per window index (int64); the batch itself never crosses PCIe

    def stage(self, idx: torch.Tensor, dtype):
        x = ops.window_stage(self.series, idx, self.seq_len, dtype, norm=self.norm, month=self.month)
        y = self.targets.index_select(0, idx)                 # row gather (plumbing)
        return x, y
