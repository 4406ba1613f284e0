"""Opacities container (stardis/radiation_field/opacities/base.py:4-28)."""
from __future__ import annotations

import numpy as np

from ...device_array import DeviceArray, as_host


class Opacities:
    """``opacities_dict``: per-source opacity arrays by name; ``total_alphas``: their sum.

    Entries written by the device path are ``DeviceArray`` objects (converted to numpy on first touch); the
    container itself behaves like the reference's."""

    def __init__(self, frequencies, stellar_model):
        self.opacities_dict = {}
        self._shape = (stellar_model.no_of_depth_points, len(frequencies))
        self._total = None

    @property
    def total_alphas(self):
        if self._total is None:
            self._total = np.zeros(self._shape)
        return self._total

    @total_alphas.setter
    def total_alphas(self, value):
        self._total = value

    def calc_total_alphas(self):
        """opacities/base.py:24-28: add every entry whose key names neither gammas nor Doppler widths to the
        (accumulating) total.  Host-side restatement for dictionaries filled by hand; ``calc_alphas`` computes the
        same sum on the device in the fused continuum pass."""
        total = as_host(self.total_alphas)
        if isinstance(self._total, DeviceArray):
            total = np.array(total)
        for key, value in self.opacities_dict.items():
            if "gammas" not in key and "doppler" not in key:
                total += as_host(value)
        self._total = total
        return self._total
