"""Loader for the UNMODIFIED reference hot-path modules (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package ``stardis_b200``.

The reference (``/root/reference``, tardis-sn/stardis) is pure Python + numba but
cannot be imported as a package in this container because ``astropy`` and
``tardis`` are absent.  Its hot-path modules, however, only use a handful of
scalar constants from ``astropy.constants`` and one helper from
``tardis.util.base`` at import time.  This module places tiny stand-ins for
those in ``sys.modules`` and then executes the reference's own source files,
from where they lie, under their real dotted names.

It is used ONLY
  * by ``oracle/make_golden*.py`` to generate the fixtures in ``tests/golden/``
    (in the build container, where ``/root/reference`` exists), and
  * by ``oracle/ref_leg.py``, the CPU reference leg of ``bench.py`` (on the GPU
    box the modules come from ``oracle/_ref``, staged verbatim and git-ignored
    by ``oracle/stage_ref.py``).

Nothing of the reference enters this repository's history.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/stage_ref.py (travels to the GPU box)
REFERENCE_ROOT = os.environ.get(
    "STARDIS_REFERENCE_ROOT",
    "/root/reference" if os.path.isdir("/root/reference/stardis/radiation_field") else _STAGED)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stardis", "radiation_field"))


class _Q(float):
    """float that tolerates the astropy attribute chain used at module level."""

    cgs = property(lambda self: self)
    esu = property(lambda self: self)
    gauss = property(lambda self: self)
    si = property(lambda self: self)
    value = property(lambda self: float(self))

    def to(self, *a, **k):
        return self

    def _w(name):
        def op(self, *args):
            r = getattr(float, name)(self, *[float(a) if isinstance(a, _Q) else a for a in args])
            return _Q(r) if isinstance(r, float) else r

        return op

    for _n in (
        "__add__ __radd__ __sub__ __rsub__ __mul__ __rmul__ __truediv__ __rtruediv__ "
        "__pow__ __rpow__ __neg__"
    ).split():
        locals()[_n] = _w(_n)
    del _n, _w


class Unit:
    """Stand-in for an astropy unit: only what the reference's hot-path modules touch."""

    __array_ufunc__ = None  # let ndarray * Unit fall through to __rmul__

    def __init__(self, name):
        self.name = name

    def __rmul__(self, other):
        return UQ(other, self)

    def __mul__(self, other):
        return Unit(f"{self.name}*{getattr(other, 'name', other)}")

    def __truediv__(self, other):
        return Unit(f"{self.name}/{getattr(other, 'name', other)}")

    def __pow__(self, p):
        return Unit(f"{self.name}**{p}")


class UQ:
    """value-with-unit: supports .value, .to(nm->AA, Hz->AA spectral), comparisons against arrays."""

    __array_ufunc__ = None

    def __init__(self, value, unit):
        self.v = value
        self.unit = unit

    value = property(lambda self: self.v)
    cgs = property(lambda self: self)

    def to(self, unit, equivalencies=None):
        src, dst = self.unit.name, unit.name
        if src == dst:
            return UQ(self.v, unit)
        if src == "nm" and dst in ("AA", "Angstrom"):
            return UQ(np.asarray(self.v) * 10.0, unit)
        if src == "Hz" and dst in ("AA", "Angstrom"):
            return UQ(_CONSTANTS["c"] / np.asarray(self.v, dtype=float) * 1e8, unit)
        if src in ("AA", "Angstrom") and dst == "Hz":
            return UQ(_CONSTANTS["c"] / (np.asarray(self.v, dtype=float) * 1e-8), unit)
        if src == "km/s" and dst == "cm/s":
            return UQ(np.asarray(self.v) * 1e5, unit)
        raise NotImplementedError((src, dst))

    def __truediv__(self, other):
        return UQ(self.v, self.unit / other) if isinstance(other, Unit) else UQ(self.v / other, self.unit)

    def __mul__(self, other):
        return UQ(self.v, self.unit * other) if isinstance(other, Unit) else UQ(self.v * other, self.unit)

    __rmul__ = __mul__

    def __lt__(self, other):  # reflected form of  array > UQ
        return np.asarray(other) > self.v

    def __gt__(self, other):
        return np.asarray(other) < self.v


class FakeQuantity(np.ndarray):
    """ndarray with the few astropy.Quantity affordances the reference's continuum functions use."""

    def __new__(cls, arr, unit="Hz"):
        obj = np.array(arr, dtype=np.float64).view(cls)
        obj.unit_name = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit_name = getattr(obj, "unit_name", "Hz")

    @property
    def value(self):
        return self.view(np.ndarray)

    @property
    def cgs(self):
        return self

    def to(self, unit, equivalencies=None):
        return UQ(self.view(np.ndarray), Unit(self.unit_name)).to(unit, equivalencies)

    def __getitem__(self, key):
        r = np.ndarray.__getitem__(self, key)
        if not isinstance(r, np.ndarray):
            return _Q(r)
        return r


def species_string_to_tuple(s):
    """'H I' -> (1, 0).  Own minimal stand-in for tardis.util.base.species_string_to_tuple."""
    sym = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K",
           "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn"]
    roman = ["I", "II", "III", "IV", "V", "VI", "VII", "VIII", "IX", "X"]
    el, ion = s.split()
    return sym.index(el) + 1, roman.index(ion)


# CODATA-2018 CGS values = astropy 6.1 (the version the reference's lock files pin).
_CONSTANTS = dict(
    c=2.99792458e10,
    h=6.62607015e-27,
    k_B=1.380649e-16,
    e=4.803204712570263e-10,
    m_e=9.1093837015e-28,
    m_p=1.67262192369e-24,
    u=1.66053906660e-24,
    a0=5.29177210903e-9,
    Ryd=109737.31568160,
    sigma_T=6.6524587321e-25,
)

_LOADED = {}


def _install_shims():
    if "astropy" in sys.modules and not getattr(sys.modules["astropy"], "_stardis_b200_shim", False):
        return  # a real astropy is present: use it
    astropy = types.ModuleType("astropy")
    astropy._stardis_b200_shim = True
    astropy.__path__ = []
    consts = types.ModuleType("astropy.constants")

    class _C(_Q):
        # const.c.to(u.km / u.s) is evaluated at import of broadening.py (C_KMS)
        def to(self, *a, **k):
            return _Q(float(self) * 1e-5)

    for k, v in _CONSTANTS.items():
        setattr(consts, k, _C(v) if k == "c" else _Q(v))
    units = types.ModuleType("astropy.units")
    for name in "km s Hz AA cm K eV erg Angstrom nm".split():
        setattr(units, name, Unit(name))
    units.spectral = lambda: "spectral"
    astropy.constants = consts
    astropy.units = units
    sys.modules["astropy"] = astropy
    sys.modules["astropy.constants"] = consts
    sys.modules["astropy.units"] = units

    if "tardis" not in sys.modules:
        tardis = types.ModuleType("tardis")
        tardis.__path__ = []
        tu = types.ModuleType("tardis.util")
        tu.__path__ = []
        tub = types.ModuleType("tardis.util.base")

        tub.species_string_to_tuple = species_string_to_tuple
        sys.modules["tardis"] = tardis
        sys.modules["tardis.util"] = tu
        sys.modules["tardis.util.base"] = tub

    for pkg in (
        "stardis",
        "stardis.radiation_field",
        "stardis.radiation_field.opacities",
        "stardis.radiation_field.opacities.opacities_solvers",
        "stardis.radiation_field.source_functions",
        "stardis.radiation_field.radiation_field_solvers",
    ):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m


def _load(dotted, relpath):
    if dotted in _LOADED:
        return _LOADED[dotted]
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    _LOADED[dotted] = mod
    return mod


class Reference:
    """Namespace with the reference's hot-path modules, loaded unmodified."""

    def __init__(self):
        if not reference_available():
            raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
        _install_shims()
        base = "stardis/radiation_field/"
        osol = "stardis.radiation_field.opacities.opacities_solvers"
        self.voigt = _load(osol + ".voigt", base + "opacities/opacities_solvers/voigt.py")
        self.broadening = _load(osol + ".broadening", base + "opacities/opacities_solvers/broadening.py")
        self.util = _load(osol + ".util", base + "opacities/opacities_solvers/util.py")
        self.opac = _load(osol + ".base", base + "opacities/opacities_solvers/base.py")
        self.blackbody = _load(
            "stardis.radiation_field.source_functions.blackbody", base + "source_functions/blackbody.py"
        )
        self.opacities_container = _load("stardis.radiation_field.opacities.base", base + "opacities/base.py")
        self.solver = _load(
            "stardis.radiation_field.radiation_field_solvers.base", base + "radiation_field_solvers/base.py"
        )


_REF = None


def load_reference() -> Reference:
    global _REF
    if _REF is None:
        _REF = Reference()
    return _REF
