"""Drop-in replacements for the reference's `src` model modules (same class names, constructor and
forward signatures, parameter names); bodies run on the pcm_b200 CUDA kernels."""
from .models import get_model  # noqa: F401
