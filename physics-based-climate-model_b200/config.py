"""Process-wide compute settings for the pcm_b200 kernels."""
import torch

_COMPUTE_DTYPE = torch.bfloat16


def set_compute_dtype(dtype: torch.dtype) -> None:
    """torch.bfloat16 (default: bf16 storage/operands, fp32 accumulation and statistics) or
    torch.float32 (everything fp32 — the tight-parity path used by the tests)."""
    global _COMPUTE_DTYPE
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError(f"unsupported compute dtype {dtype}")
    _COMPUTE_DTYPE = dtype


def compute_dtype() -> torch.dtype:
    return _COMPUTE_DTYPE
