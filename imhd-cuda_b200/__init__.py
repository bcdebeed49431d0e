"""imhd-cuda_b200: B200-native (sm_100a) replacement for the Lax-Wendroff hot path of
russellmatt66/imhd-CUDA (lib/on-device: predictor, corrector, diffusion, boundary kernels, IDX3D
layout, and the host time loop of src/on-device/{main,no_diffusion}.cu).

The directory name carries a hyphen (it follows the reference's repo name), so import it with
``importlib.import_module("imhd-cuda_b200")``.  Everything computes in ``libimhd_b200.so``
(hand-written CUDA behind the C ABI of include/imhd_b200.h); there is no fallback path.
"""
from . import _lib, ops  # noqa: F401
from ._lib import PATH_A, PATH_B, ImhdError, Slab  # noqa: F401
from .ops import Context  # noqa: F401

__all__ = ["ops", "Context", "PATH_A", "PATH_B", "Slab", "ImhdError"]
