"""MARCS model atmosphere reader (own implementation; behaviour of stardis/io/model/marcs.py:20-379).

Reads the header metadata, the logarithmic abundance block and the two 56-row structure tables of a MARCS ``.mod``
file (plane-parallel or spherical, optionally gzipped) and converts them to a ``StellarModel`` with all depth
arrays flipped to run from the deepest point to the surface (marcs.py:44-46, 203-205).
"""
from __future__ import annotations

import gzip
import re
from dataclasses import dataclass

import numpy as np
import pandas as pd

from ... import units as u
from ...constants import AMU_CGS
from ...model.base import Composition, Radial1DGeometry, StellarModel

_FLOAT = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[EeDd][-+]?\d+)?"

# standard atomic weights [u], Z = 1..92 (IUPAC abridged; used when no atom data object supplies masses)
ATOMIC_WEIGHTS = np.array([
    1.008, 4.0026, 6.94, 9.0122, 10.81, 12.011, 14.007, 15.999, 18.998, 20.180, 22.990, 24.305, 26.982, 28.085,
    30.974, 32.06, 35.45, 39.948, 39.098, 40.078, 44.956, 47.867, 50.942, 51.996, 54.938, 55.845, 58.933, 58.693,
    63.546, 65.38, 69.723, 72.630, 74.922, 78.971, 79.904, 83.798, 85.468, 87.62, 88.906, 91.224, 92.906, 95.95,
    97.0, 101.07, 102.91, 106.42, 107.87, 112.41, 114.82, 118.71, 121.76, 127.60, 126.90, 131.29, 132.91, 137.33,
    138.91, 140.12, 140.91, 144.24, 145.0, 150.36, 151.96, 157.25, 158.93, 162.50, 164.93, 167.26, 168.93, 173.05,
    174.97, 178.49, 180.95, 183.84, 186.21, 190.23, 192.22, 195.08, 196.97, 200.59, 204.38, 207.2, 208.98, 209.0,
    210.0, 222.0, 223.0, 226.0, 227.0, 232.04, 231.04, 238.03])


@dataclass
class MARCSModel:
    metadata: dict
    data: pd.DataFrame
    spherical: bool
    log_abundances: np.ndarray  # A(X), H = 12, index 0 = hydrogen; -99 = absent

    def to_geometry(self):
        r = -self.data["depth"].values[::-1].astype(np.float64)
        reference_r = None
        if self.spherical:
            radius = float(u.cgs_values_of(self.metadata["radius"]))
            r = r + radius
            reference_r = radius
        return Radial1DGeometry(u.Quantity(r, u.cm), reference_r)

    def to_composition(self, atom_data=None, final_atomic_number=92):
        n = int(min(final_atomic_number, len(self.log_abundances)))
        a = self.log_abundances[:n]
        present = a > -90
        number = np.where(present, 10.0 ** (a - 12.0), 0.0)
        masses_u = _masses_from_atom_data(atom_data, n)
        mass_frac = number * masses_u
        mass_frac = mass_frac / mass_frac.sum()
        D = len(self.data)
        idx = pd.Index(np.arange(1, n + 1), name="atomic_number")
        emf = pd.DataFrame(np.repeat(mass_frac[:, None], D, axis=1), index=idx, columns=range(D))
        density = self.data["density"].values[::-1].astype(np.float64)
        nuclide_masses = pd.Series(masses_u * AMU_CGS, index=idx)
        return Composition(u.Quantity(density, "g/cm3"), emf, nuclide_masses)

    def to_stellar_model(self, atom_data=None, final_atomic_number=92, composition_source="from_model",
                         helium_mass_frac_Y=-99.0, heavy_metal_mass_frac_Z=-99.0):
        if composition_source != "from_model":
            raise NotImplementedError(
                "composition_source other than 'from_model' needs the Asplund tables of the reference's IO layer "
                "(out of the hot-path scope)")
        temperatures = u.Quantity(self.data["t"].values[::-1].astype(np.float64), u.K)
        return StellarModel(temperatures, self.to_geometry(), self.to_composition(atom_data, final_atomic_number),
                            spherical=self.spherical, microturbulence=self.metadata["microturbulence"])


def _masses_from_atom_data(atom_data, n):
    if atom_data is not None and hasattr(atom_data, "atom_data"):
        try:
            return np.asarray(atom_data.atom_data.mass.values[:n], dtype=np.float64) / AMU_CGS
        except Exception:
            pass
    return ATOMIC_WEIGHTS[:n].copy()


def _open(fpath, gzipped):
    return gzip.open(fpath, "rt") if gzipped else open(fpath, "rt")


def read_marcs_model(fpath, gzipped=False):
    """Parse a MARCS ``.mod`` file into a ``MARCSModel`` (stardis/io/model/marcs.py:355-379)."""
    fpath = str(fpath)
    if fpath.endswith(".gz"):
        gzipped = True
    with _open(fpath, gzipped) as fh:
        lines = fh.read().splitlines()
    first = lambda i: float(re.search(_FLOAT, lines[i]).group(0).replace("D", "E"))  # noqa: E731
    spherical = "plane-parallel" not in lines[5]
    meta = {"fname": lines[0].strip()}
    meta["teff"] = u.Quantity(first(1), u.K)
    meta["flux"] = first(2)
    meta["surface_grav"] = first(3)
    meta["microturbulence"] = u.Quantity(first(4), u.km_s)
    meta["mass"] = first(5)
    feh = re.findall(_FLOAT, lines[6])
    meta["feh"], meta["afe"] = float(feh[0]), float(feh[1])
    meta["radius"] = u.Quantity(first(7), u.cm)
    meta["luminosity"] = first(8)
    conv = re.findall(_FLOAT, lines[9])
    meta["conv_alpha"], meta["conv_nu"], meta["conv_y"], meta["conv_beta"] = (float(v) for v in conv[:4])
    xyz = re.findall(_FLOAT, lines[10])
    meta["x"], meta["y"], meta["z"] = (float(v) for v in xyz[:3])

    i_ab = next(i for i, ln in enumerate(lines) if ln.startswith("Logarithmic chemical number abundances"))
    i_nd = next(i for i, ln in enumerate(lines) if "Number of depth points" in ln)
    abund = np.array([float(v) for ln in lines[i_ab + 1:i_nd] for v in ln.split()])
    n_depth = int(lines[i_nd].split()[0])

    def table(header_start):
        i0 = next(i for i, ln in enumerate(lines) if ln.split()[:len(header_start)] == header_start)
        cols = [c.lower() for c in lines[i0].split()]
        rows = []
        for ln in lines[i0 + 1:i0 + 1 + n_depth]:
            ln = re.sub(r"(?<=\d)(?=-\d)", " ", ln) if len(ln.split()) < len(cols) else ln  # fused negative columns
            rows.append([float(v) for v in ln.split()])
        return pd.DataFrame(rows, columns=cols)

    upper = table(["k", "lgTauR", "lgTau5"])
    lower = table(["k", "lgTauR", "KappaRoss"])
    data = pd.merge(upper, lower, on=["k", "lgtaur"]).set_index("k")
    return MARCSModel(meta, data, spherical, abund)
