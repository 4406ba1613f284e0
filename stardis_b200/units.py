"""Minimal unit handling for the drop-in API when astropy is not installed.

The reference passes ``astropy.units.Quantity`` objects across its API (``run_stardis(config, lambdas * u.AA)``,
``stellar_model.temperatures``, ``STARDISOutput.spectrum_lambda``).  This module provides the small subset the
hot path needs -- ``Quantity(value, unit)`` with ``.value``, ``.unit``, ``.to(unit, equivalencies)``, ``.cgs`` and
basic arithmetic -- for Angstrom/nm/cm/Hz/K/km/s/erg.  Real astropy quantities are accepted everywhere a Quantity
is expected (duck typing on ``.to`` / ``.value``).
"""
from __future__ import annotations

import numpy as np

from .constants import C_CGS

# unit name -> (dimension, factor to cgs)
_UNITS = {
    "AA": ("length", 1e-8), "Angstrom": ("length", 1e-8), "nm": ("length", 1e-7), "um": ("length", 1e-4),
    "cm": ("length", 1.0), "m": ("length", 100.0), "km": ("length", 1e5),
    "Hz": ("frequency", 1.0), "K": ("temperature", 1.0), "s": ("time", 1.0), "g": ("mass", 1.0),
    "erg": ("energy", 1.0), "eV": ("energy", 1.602176634e-12),
    "km/s": ("velocity", 1e5), "cm/s": ("velocity", 1.0),
    "g/cm3": ("density", 1.0), "1/cm3": ("number_density", 1.0),
    "erg/s/cm2/Hz": ("flux_nu", 1.0), "erg/s/cm2/AA": ("flux_lambda", 1.0),
    "": ("dimensionless", 1.0),
}


class Unit:
    def __init__(self, name):
        if name not in _UNITS:
            raise ValueError(f"unknown unit {name!r}")
        self.name = name
        self.dimension, self.cgs_factor = _UNITS[name]

    def __rmul__(self, value):
        return Quantity(value, self)

    def __truediv__(self, other):
        return Unit(f"{self.name}/{other.name}")

    def __eq__(self, other):
        return isinstance(other, Unit) and other.name == self.name

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return f"Unit({self.name})"


def spectral():
    """Marker mirroring ``astropy.units.spectral()`` (wavelength <-> frequency equivalence)."""
    return "spectral"


class Quantity(np.ndarray):
    """ndarray carrying a unit; mirrors the slice of astropy's Quantity used by the STARDIS API."""

    __array_priority__ = 100

    def __new__(cls, value, unit=""):
        obj = np.array(getattr(value, "value", value), dtype=np.float64, copy=True).view(cls)
        obj.unit = unit if isinstance(unit, Unit) else Unit(unit)
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", Unit(""))

    @property
    def value(self):
        return self.view(np.ndarray)

    @property
    def cgs(self):
        return Quantity(self.value * self.unit.cgs_factor, _cgs_unit(self.unit.dimension))

    def to(self, unit, equivalencies=None):
        unit = unit if isinstance(unit, Unit) else Unit(unit)
        src = self.unit
        if src.dimension == unit.dimension:
            return Quantity(self.value * (src.cgs_factor / unit.cgs_factor), unit)
        if {src.dimension, unit.dimension} == {"length", "frequency"}:
            if equivalencies is None:
                raise ValueError("length <-> frequency needs equivalencies=spectral()")
            cgs = self.value * src.cgs_factor
            with np.errstate(divide="ignore"):
                return Quantity((C_CGS / cgs) / unit.cgs_factor, unit)
        raise ValueError(f"cannot convert {src.name} to {unit.name}")

    def __getitem__(self, key):
        r = np.ndarray.__getitem__(self, key)
        if not isinstance(r, np.ndarray):
            return Quantity(r, self.unit)
        return r

    def __repr__(self):
        return f"<Quantity {self.value!r} {self.unit.name}>"


def _cgs_unit(dimension):
    return {"length": "cm", "frequency": "Hz", "temperature": "K", "time": "s", "mass": "g", "energy": "erg",
            "velocity": "cm/s", "density": "g/cm3", "number_density": "1/cm3", "flux_nu": "erg/s/cm2/Hz",
            "flux_lambda": "erg/s/cm2/AA", "dimensionless": ""}[dimension]


AA = Angstrom = Unit("AA")
nm = Unit("nm")
cm = Unit("cm")
km = Unit("km")
Hz = Unit("Hz")
K = Unit("K")
s = Unit("s")
g = Unit("g")
erg = Unit("erg")
eV = Unit("eV")
km_s = Unit("km/s")
cm_s = Unit("cm/s")


def to_hz(q):
    """tracing_lambdas_or_nus.to(u.Hz, u.spectral()) for our Quantity or a real astropy Quantity."""
    if isinstance(q, Quantity):
        return q.to(Hz, spectral())
    if hasattr(q, "to") and hasattr(q, "unit"):  # astropy
        import astropy.units as au

        return Quantity(q.to(au.Hz, au.spectral()).value, Hz)
    raise TypeError("tracing_lambdas_or_nus must carry units (stardis_b200.units.Quantity or astropy Quantity)")


def values_of(q):
    """Plain float64 ndarray of anything quantity-like (cgs for our Quantity, .value otherwise)."""
    if isinstance(q, Quantity):
        return q.value
    if hasattr(q, "value"):
        return np.asarray(q.value, dtype=np.float64)
    return np.asarray(q, dtype=np.float64)


def cgs_values_of(q):
    if isinstance(q, Quantity):
        return q.cgs.value
    if hasattr(q, "cgs"):
        return np.asarray(q.cgs.value, dtype=np.float64)
    return np.asarray(q, dtype=np.float64)
