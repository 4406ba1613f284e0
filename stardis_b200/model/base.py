"""Plain containers mirroring stardis/model/base.py:5-49 and stardis/model/geometry/radial1d.py:1-24."""
from __future__ import annotations

import numpy as np

from .. import units as u


class Radial1DGeometry:
    """r: radial coordinate of the depth points [cm], deepest first (model/geometry/radial1d.py)."""

    def __init__(self, r, reference_r=None):
        self.r = r
        self.reference_r = reference_r

    @property
    def dist_to_next_depth_point(self):
        r = u.values_of(self.r)
        return r[1:] - r[:-1]


class Composition:
    """The slice of tardis' Composition the hot path reads: ``nuclide_masses`` (pandas Series Z -> grams) and the
    density / elemental mass fractions used by plasma providers."""

    def __init__(self, density, elemental_mass_fraction, nuclide_masses):
        self.density = density
        self.elemental_mass_fraction = elemental_mass_fraction
        self.nuclide_masses = nuclide_masses


class StellarModel:
    """stardis/model/base.py:5-49: temperatures (deepest -> surface), geometry, composition, spherical flag,
    microturbulence."""

    hdf_properties = ["temperatures", "geometry", "composition"]

    def __init__(self, temperatures, geometry, composition, spherical=False, microturbulence=None):
        self.temperatures = temperatures
        self.geometry = geometry
        self.composition = composition
        self.spherical = spherical
        self.microturbulence = u.Quantity(0.0, u.km_s) if microturbulence is None else microturbulence

    @property
    def no_of_depth_points(self):
        return np.shape(self.temperatures)[0]
