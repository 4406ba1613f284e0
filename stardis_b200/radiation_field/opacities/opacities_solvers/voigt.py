"""Faddeeva / Voigt profile (stardis/radiation_field/opacities/opacities_solvers/voigt.py), device twins.

The reference ships ``faddeeva``/``voigt_profile`` (numba CPU ufuncs) and ``faddeeva_cuda``/``voigt_profile_cuda``
(numba.cuda wrappers that are not wired up, voigt.py:94-110, 158-195).  Here all four names run the hand-written
kernels of libstardis_b200.so (Humlicek W4, identical region logic; csrc/sd_math.cuh)."""
from __future__ import annotations

import numpy as np

from ....device import default_context

SQRT_PI = float(np.sqrt(np.pi))
PI = float(np.pi)


def faddeeva(z):
    """w(z), complex128, elementwise (voigt.py:17-91)."""
    z = np.asarray(z, dtype=np.complex128)
    out = default_context().faddeeva(np.atleast_1d(z))
    return out.reshape(z.shape) if z.ndim else complex(out[0])


def voigt_profile(delta_nu, doppler_width, gamma):
    """phi = Re w((delta_nu + i gamma/(sqrt(pi) pi)) / doppler_width) / (sqrt(pi) doppler_width) (voigt.py:113-155).
    A zero Doppler width raises ZeroDivisionError like the reference (test_voigt.py:130-148)."""
    dn, dw, g = np.broadcast_arrays(np.asarray(delta_nu, dtype=np.float64), np.asarray(doppler_width, dtype=np.float64),
                                    np.asarray(gamma, dtype=np.float64))
    if np.any(dw == 0):
        raise ZeroDivisionError("division by zero")
    out = default_context().voigt_profile(np.atleast_1d(dn), np.atleast_1d(dw), np.atleast_1d(g))
    return out.reshape(dn.shape) if dn.ndim else float(out[0])


def faddeeva_cuda(z, nthreads=256, ret_np_ndarray=True):
    """voigt.py:103-110.  ``nthreads`` is accepted for signature compatibility; results are numpy arrays."""
    return faddeeva(z)


def voigt_profile_cuda(delta_nu, doppler_width, gamma, nthreads=256, ret_np_ndarray=True):
    """voigt.py:168-195."""
    return voigt_profile(delta_nu, doppler_width, gamma)
