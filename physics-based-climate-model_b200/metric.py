"""cos(lat)-area-weighted metric on the GPU — interface of the reference's metric path:
`get_lat_weights` (src/utils_final.py:387-406), `calculate_weighted_metric` (:282-302), the triplet
of main_final.py:616-631 and the final weighting of _climate_kaggle_metric.py:109-115,144-153.

Predictions/targets stay on the device (no per-batch D2H as in main_final.py:570-571); one pass
over the data produces per-pixel time sums in fp64, a second tiny kernel does the latitude-weighted
warp-shuffle reduction.  Time-sharded inputs (data-parallel validation) add their partial sums."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib

VAR_WEIGHTS = {"tas": 0.5, "pr": 0.5}
METRIC_VAR_WEIGHTS = {
    "tas": {"monthly_rmse": 0.1, "time_mean": 1.0, "time_std": 1.0},
    "pr": {"monthly_rmse": 0.1, "time_mean": 1.0, "time_std": 0.75},
}


def get_lat_weights(latitude_values) -> np.ndarray:
    """cos(lat) normalised to mean 1 (host side: Y values)."""
    w = np.cos(np.deg2rad(np.asarray(latitude_values, dtype=np.float64)))
    return w / np.mean(w)


def _stream():
    return torch.cuda.current_stream().cuda_stream


PARTIAL_WIDTH = 8      # per pixel: sums of p, p^2, t, t^2, (p-t)^2 and the counts of non-NaN p, t, (p-t)


def metric_partial_sums(pred: torch.Tensor, truth: torch.Tensor, partial: torch.Tensor = None) -> torch.Tensor:
    """pred/truth (T, V, Y, X) fp32 (or fp64) CUDA -> fp64 (V, Y, X, 8) per-pixel sums over time of
    p, p^2, t, t^2, (p-t)^2 and the non-NaN counts (NaNs are skipped like xarray's reductions,
    src/utils_final.py:296).  Pass `partial` to accumulate further time shards into it."""
    if not pred.is_cuda:
        raise RuntimeError("pcm_b200.metric: tensors must live on the GPU (no CPU fallback)")
    f64 = pred.dtype == torch.float64
    pred = pred.contiguous() if f64 else pred.contiguous().float()
    truth = truth.contiguous().to(pred.dtype)
    T, V, Y, X = pred.shape
    zero = partial is None
    if zero:
        partial = torch.empty((V, Y, X, PARTIAL_WIDTH), device=pred.device, dtype=torch.float64)
    lib().call("pcm_metric_partial_f64" if f64 else "pcm_metric_partial", pred.data_ptr(), truth.data_ptr(),
               partial.data_ptr(), T, V, Y, X, int(zero), _stream())
    return partial


_KINDS = {"zscore": 0, "minimax": 0, "log1p": 1, "sqrt": 2, "pow": 3}


def transform_table(output_stats: dict, n_vars: int, device) -> torch.Tensor:
    """Normalizer.output_stats ({var_idx: {"method": ..., "params": {...}}}, src/utils_final.py:130-206) -> the
    device table (V, 4) = (kind, a, b, c) pcm_metric_partial_denorm takes; variables without an entry pass through."""
    rows = []
    for v in range(n_vars):
        cfg = output_stats.get(v)
        if cfg is None:
            rows.append([0.0, 1.0, 0.0, 1.0])
            continue
        m, pr = cfg["method"], cfg.get("params", {})
        if m not in _KINDS:
            raise ValueError(f"Unknown inverse method '{m}' for var {v}.")
        if m == "minimax":
            a, b = pr["max_val"] - pr["min_val"], pr["min_val"]
        else:
            a, b = pr["std"], pr["mean"]
        rows.append([float(_KINDS[m]), float(a), float(b), float(pr.get("lambda", 1.0))])
    return torch.tensor(rows, dtype=torch.float32, device=device)


def metric_partial_sums_normalized(pred_norm: torch.Tensor, truth_norm: torch.Tensor, table: torch.Tensor,
                                   partial: torch.Tensor = None) -> torch.Tensor:
    """As metric_partial_sums, for NORMALISED predictions/targets: the inverse transform of the reference's Normalizer is
    applied on the fly (validation epilogue of main_final.py:563-574 without the de-normalised copies or the D2H)."""
    if not pred_norm.is_cuda:
        raise RuntimeError("pcm_b200.metric: tensors must live on the GPU (no CPU fallback)")
    pred_norm = pred_norm.contiguous().float()
    truth_norm = truth_norm.contiguous().float()
    T, V, Y, X = pred_norm.shape
    zero = partial is None
    if zero:
        partial = torch.empty((V, Y, X, PARTIAL_WIDTH), device=pred_norm.device, dtype=torch.float64)
    lib().call("pcm_metric_partial_denorm", pred_norm.data_ptr(), truth_norm.data_ptr(), table.data_ptr(), partial.data_ptr(),
               T, V, Y, X, int(zero), _stream())
    return partial


def metric_finalize(partial: torch.Tensor, lat, T_total: int, weights=None) -> torch.Tensor:
    """-> fp64 (V, 3): monthly_rmse, time_mean_rmse, time_std_mae per variable.  `weights` (Y,) overrides the
    cos(lat)/mean weights (the Kaggle scorer passes cos(lat_2dp)/sum; any positive scaling gives the same result)."""
    V, Y, X, _ = partial.shape
    w = torch.as_tensor(get_lat_weights(lat) if weights is None else np.asarray(weights, np.float64),
                        dtype=torch.float64, device=partial.device)
    assert w.numel() == Y
    out = torch.empty((V, 3), device=partial.device, dtype=torch.float64)
    lib().call("pcm_metric_finalize", partial.data_ptr(), w.data_ptr(), out.data_ptr(), int(T_total), V, Y, X, _stream())
    return out


def weighted_metric_triplets(pred: torch.Tensor, truth: torch.Tensor, lat):
    """[(monthly_rmse, time_mean_rmse, time_std_mae)] per variable (python floats; one D2H of V*3 doubles)."""
    part = metric_partial_sums(pred, truth)
    out = metric_finalize(part, lat, pred.shape[0]).cpu().numpy()
    return [tuple(float(v) for v in row) for row in out]


def combined_score(triplets: dict) -> float:
    """{var: (monthly, tmean, tstd)} -> competition score (lower is better)."""
    s = 0.0
    for var, (m, tm, ts) in triplets.items():
        k = METRIC_VAR_WEIGHTS[var]
        s += VAR_WEIGHTS[var] * (k["monthly_rmse"] * m + k["time_mean"] * tm + k["time_std"] * ts)
    return float(s)
