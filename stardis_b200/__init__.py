"""stardis_b200 -- B200-native implementation of the STARDIS opacity + formal-solution hot path.

Drop-in for ``stardis``' ``run_stardis`` / ``calc_alphas`` / ``raytrace`` / ``RadiationField`` / ``STARDISOutput``
(same names, arguments and side effects); the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``libstardis_b200.so`` (include/stardis_b200.h).  No CPU fallback.
"""
from .base import *  # noqa: F401,F403
from .base import run_stardis, set_num_threads, STARDISOutput  # noqa: F401

__version__ = "0.1.0"
