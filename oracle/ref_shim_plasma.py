"""Loader for the UNMODIFIED reference module ``stardis/plasma/base.py`` (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package ``stardis_b200``.

``AlphaLineVald`` / ``AlphaLineShortlistVald`` (plasma/base.py:178-455) use astropy for real unit algebra
(``values * u.eV``, ``.to(1)``, ``.to(u.Hz, equivalencies=u.spectral())``), so the name-based stand-ins of
``ref_shim.py`` are not enough.  Here a unit is a CGS scale factor and a quantity is an ndarray subclass holding CGS
numbers: products and quotients are then plain numpy arithmetic, ``.to(1)`` / ``.cgs`` are the identity, a spectral
conversion is ``c / x`` and ``.to(unit).value`` divides the scale out again.  The tardis base classes the module
derives from are empty stand-ins.  Must run in its own process (it installs different ``astropy`` stand-ins than
``ref_shim``); used only by ``oracle/make_golden_plasma.py`` in the build container.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("STARDIS_REFERENCE_ROOT", "/root/reference")
C_CGS, H_CGS, KB_CGS = 2.99792458e10, 6.62607015e-27, 1.380649e-16
EV_ERG = 1.602176634e-12


class CGS(np.ndarray):
    """CGS numbers with the few Quantity methods the plasma module calls."""

    def __new__(cls, arr):
        return np.asarray(arr, dtype=np.float64).view(cls)

    @property
    def value(self):
        return np.asarray(self)

    cgs = property(lambda self: self)

    def __array_function__(self, func, types, args, kwargs):  # np.outer & co. keep the Quantity type, as astropy does
        plain = [np.asarray(a) if isinstance(a, CGS) else a for a in args]
        r = func(*plain, **kwargs)
        return CGS(r) if isinstance(r, np.ndarray) else r

    def to(self, unit, equivalencies=None):
        if equivalencies is not None:  # spectral: wavelength [cm] <-> frequency [Hz]
            return CGS(C_CGS / np.asarray(self))
        scale = getattr(unit, "scale", 1.0)
        return CGS(np.asarray(self) / scale)


class ScaleUnit:
    __array_ufunc__ = None

    def __init__(self, scale):
        self.scale = float(scale)

    def __rmul__(self, other):
        return CGS(np.asarray(other, dtype=np.float64) * self.scale)

    def __mul__(self, other):
        if isinstance(other, ScaleUnit):
            return ScaleUnit(self.scale * other.scale)
        return CGS(np.asarray(other, dtype=np.float64) * self.scale)

    def __truediv__(self, other):
        return ScaleUnit(self.scale / other.scale)

    def __pow__(self, p):
        return ScaleUnit(self.scale ** p)


class Const(float):
    cgs = property(lambda self: self)
    gauss = property(lambda self: self)
    esu = property(lambda self: self)
    value = property(lambda self: float(self))

    def _w(name):
        def op(self, *args):
            r = getattr(float, name)(self, *[float(a) if isinstance(a, Const) else a for a in args])
            return Const(r) if isinstance(r, float) else r
        return op

    for _n in "__add__ __radd__ __sub__ __rsub__ __mul__ __rmul__ __truediv__ __rtruediv__ __pow__ __neg__".split():
        locals()[_n] = _w(_n)
    del _n, _w


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_plasma_module():
    consts = _module("astropy.constants", c=Const(C_CGS), h=Const(H_CGS), k_B=Const(KB_CGS), e=Const(4.803204712570263e-10),
                     m_e=Const(9.1093837015e-28), m_p=Const(1.67262192369e-24), u=Const(1.66053906660e-24))
    units = _module("astropy.units", eV=ScaleUnit(EV_ERG), K=ScaleUnit(1.0), AA=ScaleUnit(1e-8), Angstrom=ScaleUnit(1e-8),
                    Hz=ScaleUnit(1.0), cm=ScaleUnit(1.0), erg=ScaleUnit(1.0), s=ScaleUnit(1.0), km=ScaleUnit(1e5),
                    Pa=ScaleUnit(10.0),
                    spectral=lambda: "spectral")
    _module("astropy", constants=consts, units=units)

    class _Base:
        def __init__(self, plasma_parent=None):
            self.plasma_parent = plasma_parent

    _module("tardis")
    _module("tardis.util")
    symbols = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn").split()
    _module("tardis.util.base", element_symbol2atomic_number=lambda s: symbols.index(s) + 1,
            species_string_to_tuple=lambda s: (0, 0))
    _module("tardis.plasma")
    _module("tardis.plasma.base", BasePlasma=_Base)
    _module("tardis.plasma.properties")
    _module("tardis.plasma.properties.base", DataFrameInput=_Base, ProcessingPlasmaProperty=_Base)
    _module("tardis.plasma.properties.property_collections", basic_inputs=[], basic_properties=[],
            lte_excitation_properties=[], lte_ionization_properties=[], non_nlte_properties=[], helium_lte_properties=[])
    _module("tardis.opacities")
    _module("tardis.opacities.tau_sobolev", TauSobolev=_Base)
    _module("stardis")
    _module("stardis.plasma")
    out = {}
    for dotted, rel in (("stardis.plasma.molecules", "stardis/plasma/molecules.py"), ("stardis.plasma.base", "stardis/plasma/base.py")):
        spec = importlib.util.spec_from_file_location(dotted, os.path.join(REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[dotted] = mod
        spec.loader.exec_module(mod)
        out[dotted] = mod
    return out["stardis.plasma.base"]
