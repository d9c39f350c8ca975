"""hypernerf_torch_b200 — B200-native per-ray hot path of HyperNeRF behind the reference's Python interface.

(The directory is `hypernerf_torch_b200`, not `hypernerf-torch_b200`, so that it is importable.)

    from hypernerf_torch_b200.models import NerfModel          # drop-in for hypernerf.models.NerfModel
    from hypernerf_torch_b200 import model_utils                # drop-in for hypernerf.model_utils (hot-path part)

Everything numerical runs in libhypernerf_b200.so (hand-written sm_100a CUDA, C-ABI in include/hypernerf_b200.h);
there is no CPU or torch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "models", "model_utils", "modules"]
