"""Cross-section tables and number densities of the file/continuum opacity sources.

Host-side mirror of stardis/radiation_field/opacities/opacities_solvers/util.py:14-166.  The tables are tiny
(<= 42 x 11 numbers); they are parsed on the host, the Delaunay split scipy's ``LinearNDInterpolator`` would use is
extracted ONCE per table (one diagonal flag per rectangular cell) and the D x N interpolation itself runs on the
device inside the fused continuum kernel (csrc/k3_continuum.cu).
"""
from __future__ import annotations

import logging
from functools import lru_cache
from pathlib import Path

import numpy as np

from ....constants import KB_CGS

logger = logging.getLogger(__name__)

_SYMBOLS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K", "Ca",
            "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y",
            "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe", "Cs", "Ba", "La", "Ce",
            "Pr", "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir",
            "Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn", "Fr", "Ra", "Ac", "Th", "Pa", "U"]
_ROMAN = {"I": 0, "II": 1, "III": 2, "IV": 3, "V": 4, "VI": 5, "VII": 6, "VIII": 7, "IX": 8, "X": 9, "XI": 10, "XII": 11,
          "XIII": 12, "XIV": 13, "XV": 14, "XVI": 15, "XVII": 16, "XVIII": 17, "XIX": 18, "XX": 19}


def species_string_to_tuple(species):
    """'H I' / 'Fe 2' -> (atomic_number, ion_number); stands in for tardis.util.base.species_string_to_tuple."""
    try:
        from tardis.util.base import species_string_to_tuple as tardis_impl  # use tardis when it is installed

        return tardis_impl(species)
    except ImportError:
        pass
    element, ion = species.split()
    z = _SYMBOLS.index(element.capitalize()) + 1
    ion_number = _ROMAN[ion.upper()] if ion.upper() in _ROMAN else int(ion) - 1
    if ion_number > z:
        raise ValueError(f"Species given does not exist: ion number > atomic number ({species})")
    return z, ion_number


def _rows(fpath):
    return [ln for ln in Path(fpath).read_text().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]


def _sci(tok):
    """Stancil's compact notation '7.34-5' -> 7.34e-5 (util.py:41)."""
    return float(tok.replace("-", "e-") if "-" in tok[1:] and "e" not in tok.lower() else tok)


@lru_cache(maxsize=32)
def read_table(fpath, opacity_source):
    """Parse a cross-section file into a device-table description (dict of small host arrays).

    Hminus_bf (util.py:93-103): 1-D (wavelength [A], sigma), evaluated with np.interp semantics (end clamping).
    Hminus_ff (util.py:63-91):  2-D (wavelength [A], theta = 5040/T).
    H2plus_bf (util.py:35-62):  2-D (wavelength [nm -> A], T).
    Anything else raises ValueError exactly like the reference (util.py:105-106)."""
    fpath = str(fpath)
    if opacity_source == "Hminus_bf":
        tab = np.array([[float(v) for v in ln.split(",")] for ln in _rows(fpath)])
        return dict(kind=1, x=np.ascontiguousarray(tab[:, 0]), y=None, values=np.ascontiguousarray(tab[:, 1]), diag=None)
    if opacity_source == "Hminus_ff":
        rows = _rows(fpath)
        ys = np.array([float(h) for h in (t.strip(",") for t in rows[0].split()) if h])
        body = np.array([[float(t) for t in ln.split()] for ln in rows[1:]])
        xs, vals = body[:, 0], body[:, 1:]
    elif opacity_source == "H2plus_bf":
        rows = _rows(fpath)
        ys = np.array([float(int(float(h))) for h in rows[0].split()[1:]])
        body = [[_sci(t) for t in ln.split()] for ln in rows[1:]]
        xs = np.array([r[0] for r in body]) * 10.0  # nm -> Angstrom
        vals = np.array([r[1:] for r in body])
    else:
        raise ValueError(f"Unknown opacity_source: {opacity_source}")
    if vals.shape != (xs.size, ys.size):
        raise ValueError(f"malformed cross-section table {fpath}")
    return dict(kind=2, x=np.ascontiguousarray(xs), y=np.ascontiguousarray(ys), values=np.ascontiguousarray(vals),
                diag=delaunay_diagonals(xs, ys))


def delaunay_diagonals(xs, ys):
    """Which diagonal splits each rectangular cell in the triangulation that scipy's LinearNDInterpolator builds on
    the meshgrid points (util.py:47-57, 73-80): 0 = (0,0)-(1,1), 1 = (1,0)-(0,1).  The four corners of a rectangle are
    co-circular, so the choice is qhull's; it is read back from the triangulation itself."""
    from scipy.spatial import Delaunay

    xm, ym = np.meshgrid(xs, ys, indexing="ij")
    tri = Delaunay(np.vstack([xm.ravel(), ym.ravel()]).T)
    ny = len(ys)
    diag = np.full((len(xs) - 1, len(ys) - 1), 255, dtype=np.uint8)
    for simplex in tri.simplices:
        ij = np.array([(p // ny, p % ny) for p in simplex])
        i0, j0 = ij.min(0)
        i1, j1 = ij.max(0)
        if i1 - i0 != 1 or j1 - j0 != 1:
            raise NotImplementedError("cross-section table triangulation is not cell-aligned")
        missing = ({(0, 0), (1, 0), (0, 1), (1, 1)} - {(a - i0, b - j0) for a, b in ij}).pop()
        dg = 0 if missing in ((1, 0), (0, 1)) else 1
        if diag[i0, j0] not in (255, dg):
            raise NotImplementedError("inconsistent triangulation of a table cell")
        diag[i0, j0] = dg
    if (diag == 255).any():
        raise NotImplementedError("cross-section table triangulation does not cover every cell")
    return np.ascontiguousarray(diag)


def table_descriptor(fpath, opacity_source, temperatures, number_density):
    """Device table for calc_alpha_file (opacities_solvers/base.py:40-70): per-depth second coordinate and the
    per-depth multiplier (unit scaling x number density)."""
    t = dict(read_table(str(fpath), opacity_source))
    T = np.asarray(temperatures, dtype=np.float64)
    n = np.asarray(number_density, dtype=np.float64)
    if opacity_source == "Hminus_bf":
        t["depth_y"] = None
        t["depth_scale"] = np.ascontiguousarray(n)
    elif opacity_source == "Hminus_ff":
        t["depth_y"] = np.ascontiguousarray(5040.0 / T)
        # (interp * 1e-26 * k_B * T) * n, in the reference's order of operations (util.py:81-87, base.py:70)
        t["depth_scale"] = np.ascontiguousarray(1e-26 * KB_CGS * T * n)
    elif opacity_source == "H2plus_bf":
        t["depth_y"] = np.ascontiguousarray(T)
        t["depth_scale"] = np.ascontiguousarray(1e-18 * n)
    return t


_SINGLE_DENSITY = {"Hminus_bf": "h_minus_density", "H2plus_bf": "h2_plus_density"}
# product of two densities: "e" = electrons, "h2" = molecular hydrogen, tuple = (atomic_number, ion_number)
_PAIR_DENSITY = {"Hminus_ff": ((1, 0), "e"), "Heminus_ff": ((2, 0), "e"), "H2minus_ff": ("h2", "e"),
                 "H2plus_ff": ((1, 0), (1, 1))}


def _density(stellar_plasma, what):
    if what == "e":
        return stellar_plasma.electron_densities
    if what == "h2":
        return stellar_plasma.h2_density
    return stellar_plasma.ion_number_density.loc[what[0], what[1]]


def get_number_density(stellar_plasma, opacity_source):
    """(number_density, atomic_number, ion_number) of an opacity source string, with the semantics of util.py:111-166:
    named H-/H2+/He-/H2- sources map to fixed density products and return (density, None, None); ``<El>_<ION>_bf``
    gives n(Z, ion); ``<El>_<ION>_ff`` gives n_e * n(Z, ion + 1) and reports ion + 1 (the charge seen by the
    free electron)."""
    if opacity_source in _SINGLE_DENSITY:
        return getattr(stellar_plasma, _SINGLE_DENSITY[opacity_source]), None, None
    if opacity_source in _PAIR_DENSITY:
        first, second = _PAIR_DENSITY[opacity_source]
        return _density(stellar_plasma, first) * _density(stellar_plasma, second), None, None
    species, kind = opacity_source[:-3], opacity_source[-2:]
    atomic_number, ion_number = species_string_to_tuple(species.replace("_", " "))
    if kind == "ff":
        ion_number += 1
        return (stellar_plasma.electron_densities * stellar_plasma.ion_number_density.loc[atomic_number, ion_number],
                atomic_number, ion_number)
    return 1 * stellar_plasma.ion_number_density.loc[atomic_number, ion_number], atomic_number, ion_number
