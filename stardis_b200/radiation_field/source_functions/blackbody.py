"""Planck source function (stardis/radiation_field/source_functions/blackbody.py:11-35), evaluated on the device."""
from __future__ import annotations

import numpy as np

from ... import units as u
from ...device import default_context


def blackbody_flux_at_nu(tracing_nus, temps):
    """B_nu(T) = 2 h nu^3 / c^2 / (exp(h nu / k_B T) - 1) for nu (N,) and temperatures (D,) or (D,1) -> (D,N)
    [erg / (s cm^2 Hz)].  Runs kernel ``k_ew_blackbody`` (the formal solver evaluates the same expression inline)."""
    nus = u.values_of(tracing_nus)
    T = np.ravel(u.values_of(temps))
    return default_context().blackbody(nus, T)
