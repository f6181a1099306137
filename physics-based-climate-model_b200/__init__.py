"""pcm_b200 — B200-native (sm_100a) hot path of the Physics-Based-Climate-Model emulator.

The directory name follows the build contract (`physics-based-climate-model_b200/`); since a
hyphenated name is not importable, the repo-root shim `pcm_b200.py` registers this package as
`pcm_b200`.  Public surface:

  pcm_b200.src.*            drop-in replacements of the reference's src/ model modules
  pcm_b200.ops              autograd Functions over the C-ABI kernels (include/pcm_b200.h)
  pcm_b200.metric           cos(lat)-area-weighted metric triplet / score
  pcm_b200.optim, .trainer  flat-buffer fused Adam and the data-parallel training step
  pcm_b200.build()          compile libpcm_b200.so in-tree (nvcc, sm_100a)
"""
from . import _build
from .config import compute_dtype, set_compute_dtype  # noqa: F401


def build(force: bool = False, verbose: bool = False) -> str:
    return _build.build(force=force, verbose=verbose)


def library_path() -> str:
    return _build.LIB_PATH
