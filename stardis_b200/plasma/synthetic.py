"""Synthetic LTE-like plasma state with the attribute surface of the tardis ``BasePlasma`` that the hot path reads.

The reference's plasma (ion/level populations, ``alpha_line``) is produced by the third-party ``tardis`` package
from the Kurucz/CD23 atomic data; neither exists offline.  The hot path only READS a handful of
attributes; this module generates them with the shapes, index structure and value ranges of a real
run so that the same adapter code is exercised:

    lines, ionization_data, atomic_data.levels.energy, alpha_line, electron_densities, ion_number_density,
    level_number_density, levels, excitation_energy, h_minus_density, h2_density, h2_plus_density
    (+ lines_from_linelist / alpha_line_from_linelist with VALD stark/waals columns)

Number densities follow from the MARCS gas/electron pressures through the ideal-gas law and a hydrogen Saha
equation; line strengths follow the recipe of SURVEY.md 8(d).  It is NOT a physical plasma solve.
"""
from __future__ import annotations

import types

import numpy as np
import pandas as pd

from ..constants import ALPHA_COEFFICIENT, AMU_CGS, EV_ERG, H_CGS, KB_CGS, ME_CGS
from .columnar import ColumnarLines, LineStrength


class SyntheticPlasma:
    """Duck type of the tardis plasma for the hot path.  ``line_table`` holds the columnar, nu-sorted line data;
    the pandas views the reference reads (``lines``, ``alpha_line``...) are built lazily from it."""

    def __init__(self, T, n_e, n_HI, n_HII, n_HeI, h_minus, h2, h2_plus, line_table, h_levels=None,
                 molecule_table=None):
        D = len(T)
        cols = range(D)
        self.electron_densities = pd.Series(n_e, index=cols)
        idx = pd.MultiIndex.from_tuples([(1, 0), (1, 1), (2, 0)], names=["atomic_number", "ion_number"])
        self.ion_number_density = pd.DataFrame(np.vstack([n_HI, n_HII, n_HeI]), index=idx, columns=cols)
        self.h_minus_density = pd.Series(h_minus, index=cols)
        self.h2_density = pd.Series(h2, index=cols)
        self.h2_plus_density = pd.Series(h2_plus, index=cols)
        self._line_table = line_table
        self.expose_columnar = True  # False: only the tardis-style pandas tables are visible (adapter tests)
        self.molecule_line_table = molecule_table
        # hydrogen levels for the hydrogenic bound-free opacity
        if h_levels is None:
            h_levels = (np.zeros(0), np.zeros((0, D)))
        e_exc, n_lev = h_levels
        lidx = pd.MultiIndex.from_tuples([(1, 0, k) for k in range(len(e_exc))],
                                         names=["atomic_number", "ion_number", "level_number"])
        self.levels = lidx
        self.excitation_energy = pd.Series(e_exc, index=lidx)
        self.level_number_density = pd.DataFrame(n_lev, index=lidx, columns=cols)
        ion_rows = {(2, 1): 24.587388, (2, 2): 54.417763}
        for z, (first, second) in IONIZATION_EV.items():
            ion_rows[(z, 1)] = first
            if z > 1:
                ion_rows[(z, 2)] = second
        keys = sorted(ion_rows)
        iidx = pd.MultiIndex.from_tuples(keys, names=["atomic_number", "ion_number"])
        self.ionization_data = pd.Series(np.array([ion_rows[k] for k in keys]) * EV_ERG, index=iidx, name="ionization_energy")
        self._pandas = None
        # optional molecular line list (plasma/molecules.py attribute surface): DataFrames + molecule -> (Ion1, Ion2)
        self.molecule_lines_from_linelist = None
        self.molecule_alpha_line_from_linelist = None
        self.molecule_ion_map = None

    # ---- pandas views with the reference's layout (built on demand; used by the golden generator and the
    #      slow-path adapter test, never by the fast path) ---------------------------------------------------
    def _build_pandas(self):
        if self._pandas is not None:
            return self._pandas
        lt = self._line_table
        L = len(lt)
        # give every line its own (lower, upper) level numbers so that the reference's merges are exercised
        lower = np.arange(L) * 2
        upper = lower + 1
        idx = pd.MultiIndex.from_arrays([lt.atomic_number, lt.ion_number, lower, upper],
                                        names=["atomic_number", "ion_number", "level_number_lower", "level_number_upper"])
        lines = pd.DataFrame({"line_id": np.arange(L), "nu": lt.nu, "A_ul": lt.A_ul}, index=idx)
        lev_idx = pd.MultiIndex.from_arrays([np.repeat(lt.atomic_number, 2), np.repeat(lt.ion_number, 2),
                                             np.stack([lower, upper], 1).ravel()],
                                            names=["atomic_number", "ion_number", "level_number"])
        energy = pd.Series(np.stack([lt.level_energy_lower, lt.level_energy_upper], 1).ravel(), index=lev_idx, name="energy")
        alpha = pd.DataFrame(lt.alpha_line, index=lines.index)
        alpha["nu"] = lt.nu
        self._pandas = dict(lines=lines, levels_energy=energy, alpha=alpha)
        return self._pandas

    @property
    def line_table(self):
        return self._line_table if self.expose_columnar else None

    @property
    def lines(self):
        return self._build_pandas()["lines"]

    @property
    def alpha_line(self):
        return self._build_pandas()["alpha"]

    @property
    def atomic_data(self):
        return types.SimpleNamespace(levels=types.SimpleNamespace(energy=self._build_pandas()["levels_energy"]))

    @property
    def lines_from_linelist(self):
        lt = self._line_table
        df = pd.DataFrame({k: getattr(lt, k) for k in ("nu", "atomic_number", "ion_number", "ionization_energy",
                                                       "level_energy_lower", "level_energy_upper", "A_ul")})
        if lt.stark is not None:
            df["stark"], df["waals"] = lt.stark, lt.waals
        return df

    @property
    def alpha_line_from_linelist(self):
        a = pd.DataFrame(self._line_table.alpha_line)
        a["nu"] = self._line_table.nu
        return a


# first and second ionisation energies [eV] of the elements the synthetic line list draws from
IONIZATION_EV = {1: (13.598434, 13.598434), 6: (11.2603, 24.3833), 7: (14.5341, 29.6013), 8: (13.6181, 35.1211),
                 11: (5.1391, 47.2864), 12: (7.6462, 15.0353), 13: (5.9858, 18.8286), 14: (8.1517, 16.3459),
                 20: (6.1132, 11.8717), 22: (6.8281, 13.5755), 23: (6.7462, 14.618), 24: (6.7665, 16.4857),
                 25: (7.4340, 15.6400), 26: (7.9025, 16.1992), 27: (7.8810, 17.084), 28: (7.6399, 18.1688)}


def saha_hydrogen(T, n_e):
    """n(H II)/n(H I) from the Saha equation (partition functions 2 and 1)."""
    lam = (2 * np.pi * ME_CGS * KB_CGS * T / H_CGS**2) ** 1.5
    return lam / n_e * np.exp(-13.598434 * EV_ERG / (KB_CGS * T))


def synthetic_line_table(rng, n_lines, nu_min, nu_max, T, strong_fraction=0.005, vald=False, log_alpha=(-2.0, 8.0),
                         log_alpha_strong=(9.0, 11.0), device_strengths=False):
    """Seeded synthetic line list following SURVEY.md 8(d): uniform in nu, iron-group-heavy Z distribution,
    E_lower ~ U[0, 0.8 E_ion], E_upper = E_lower + h nu (< E_ion enforced), A_ul = 10^U[6,9], alpha_line
    log-uniform with a Boltzmann depth trend exp(-E_lower / k T_d) normalised at the hottest depth, plus a strong
    fraction whose windows span the whole grid."""
    D = T.size
    L = int(n_lines)
    nu = np.sort(rng.uniform(nu_min, nu_max, L))
    zs = np.array([1, 6, 7, 8, 11, 12, 13, 14, 20, 22, 23, 24, 25, 26, 27, 28])
    pz = np.array([2, 3, 2, 3, 2, 4, 2, 4, 5, 8, 5, 8, 5, 30, 6, 11], dtype=float)
    Z = rng.choice(zs, size=L, p=pz / pz.sum()).astype(np.int64)
    ion = np.where(Z == 1, 0, (rng.random(L) < 0.35).astype(np.int64)).astype(np.int64)
    h_nu = H_CGS * nu
    # ionisation energy is a property of the ion (the reference joins it on (atomic_number, ion_number));
    # upper level at least `gap` below the continuum so that n_eff stays in the physical range (<~ 8)
    gap = 0.25 * EV_ERG * (ion + 1.0) ** 2
    e_ion = np.array([IONIZATION_EV[z][i] for z, i in zip(Z, ion)]) * EV_ERG
    tight = e_ion - gap - h_nu <= 0.05 * EV_ERG  # photon does not fit below this ion's continuum: make it Fe II
    Z[tight], ion[tight] = 26, 1
    gap = 0.25 * EV_ERG * (ion + 1.0) ** 2
    e_ion = np.array([IONIZATION_EV[z][i] for z, i in zip(Z, ion)]) * EV_ERG
    room = e_ion - gap - h_nu
    e_lo = rng.uniform(0.0, 1.0, L) * np.minimum(room, 0.8 * e_ion)
    e_up = e_lo + h_nu
    A_ul = 10.0 ** rng.uniform(6.0, 9.0, L)
    a0 = 10.0 ** rng.uniform(log_alpha[0], log_alpha[1], L)
    strong = rng.random(L) < strong_fraction
    a0[strong] = 10.0 ** rng.uniform(log_alpha_strong[0], log_alpha_strong[1], int(strong.sum()))
    strength = None
    if device_strengths:
        # the same recipe expressed through the O(L) inputs of the reference's VALD short-list producer
        # (plasma/base.py:324-455): alpha = C (N/U)[ion, d] exp(-E_lo / k T_d) gf (1 - exp(-h nu / k T_d)) with a
        # depth-independent N/U = 1 per ion and gf chosen so that alpha = a0 at the hottest depth -- the device fills the
        # (L, D) array itself; the host copy below serves the CPU legs and the oracle only
        ions = sorted({(int(z), int(i)) for z, i in zip(Z, ion)})
        row_of = {zi: k for k, zi in enumerate(ions)}
        ion_row = np.array([row_of[(int(z), int(i))] for z, i in zip(Z, ion)], dtype=np.int64)
        t_hot = T.max()
        gf = a0 / (ALPHA_COEFFICIENT * np.exp(-e_lo * (1.0 / (t_hot * KB_CGS))) * (1.0 - np.exp((-H_CGS / KB_CGS) * (nu * (1.0 / t_hot)))))
        strength = LineStrength("vald", dict(ion_row=ion_row, gf=gf, g_lo=None), dict(n_over_u=np.ones((len(ions), D))))
        alpha = strength.host_alpha(T, nu, e_lo)
    else:
        boltz = np.exp(-e_lo[:, None] / KB_CGS * (1.0 / T[None, :] - 1.0 / T.max()))
        alpha = np.ascontiguousarray(a0[:, None] * boltz)
    stark = waals = None
    if vald:
        stark = np.where(rng.random(L) < 0.15, 0.0, -rng.uniform(4.5, 6.5, L))
        kind = rng.integers(0, 4, L)
        waals = np.where(kind == 0, -rng.uniform(7.0, 8.0, L), np.where(kind == 1, 0.0, np.where(
            kind == 2, rng.uniform(0.5, 3.0, L), rng.integers(150, 1500, L) + rng.uniform(0.15, 0.35, L))))
    return ColumnarLines(nu=nu, atomic_number=Z, ion_number=ion, ionization_energy=e_ion, level_energy_lower=e_lo,
                         level_energy_upper=e_up, A_ul=A_ul, alpha_line=alpha, stark=stark, waals=waals, strength=strength)


def attach_synthetic_molecules(plasma, T, n_lines, nu_min, nu_max, seed=0):
    """Seeded molecular line list with the attribute layout calc_molecular_alpha_line_at_nu reads
    (opacities_solvers/base.py:444-484, broadening.py:808-819)."""
    rng = np.random.default_rng(seed)
    names = np.array(["CH", "CN", "OH", "TiO", "MgH"])
    ion_map = pd.DataFrame({"Ion1": [6, 6, 8, 22, 12], "Ion2": [1, 7, 1, 8, 1]}, index=pd.Index(names, name="molecule"))
    nu = rng.uniform(nu_min, nu_max, n_lines)  # deliberately unsorted: the driver sorts by nu like the reference
    mol = names[rng.integers(0, len(names), n_lines)]
    A_ul = 10.0 ** rng.uniform(5.0, 8.0, n_lines)
    e_lo = rng.uniform(0.0, 2.0, n_lines) * EV_ERG
    a0 = 10.0 ** rng.uniform(-2.0, 4.0, n_lines)
    alpha = a0[:, None] * np.exp(-e_lo[:, None] / KB_CGS * (1.0 / T[None, :] - 1.0 / T.min()))
    plasma.molecule_lines_from_linelist = pd.DataFrame({"nu": nu, "A_ul": A_ul, "molecule": mol})
    a = pd.DataFrame(alpha)
    a["nu"] = nu
    plasma.molecule_alpha_line_from_linelist = a
    plasma.molecule_ion_map = ion_map
    return plasma


def create_synthetic_plasma(atmosphere, n_lines, nu_min, nu_max, seed=0, strong_fraction=0.005, vald=False,
                            n_h_levels=12, log_alpha=(-2.0, 8.0), log_alpha_strong=(9.0, 11.0), device_strengths=False):
    """atmosphere: dict with T, pe, pg (deepest -> surface, cgs)."""
    rng = np.random.default_rng(seed)
    T = np.asarray(atmosphere["T"], dtype=np.float64)
    pe, pg = np.asarray(atmosphere["pe"]), np.asarray(atmosphere["pg"])
    n_e = pe / (KB_CGS * T)
    n_nuclei = (pg - pe) / (KB_CGS * T)
    n_H = 0.92 * n_nuclei
    ratio = saha_hydrogen(T, n_e)
    n_HI = n_H / (1.0 + ratio)
    n_HII = n_H - n_HI
    n_HeI = 0.078 * n_nuclei
    # H- from its Saha-like equilibrium (binding energy 0.754 eV, statistical weights 1 : 2)
    lam = (2 * np.pi * ME_CGS * KB_CGS * T / H_CGS**2) ** 1.5
    h_minus = n_HI * n_e / (4.0 * lam) * np.exp(0.754 * EV_ERG / (KB_CGS * T))
    h2 = 1e-4 * n_HI * (n_HI / 1e17) * np.exp(4.478 * EV_ERG / (KB_CGS * T) - 4.478 * EV_ERG / (KB_CGS * 4000.0))
    h2_plus = 1e-9 * n_HI * n_HII / np.maximum(n_H, 1e-300) * 1e-2
    # hydrogen levels n = 1..n_h_levels (Boltzmann, g = 2 n^2)
    n = np.arange(1, n_h_levels + 1)
    e_exc = 13.598434 * EV_ERG * (1.0 - 1.0 / n**2)
    g = 2.0 * n**2
    bz = g[:, None] * np.exp(-e_exc[:, None] / (KB_CGS * T[None, :]))
    n_lev = n_HI[None, :] * bz / bz.sum(0, keepdims=True)
    lt = synthetic_line_table(rng, n_lines, nu_min, nu_max, T, strong_fraction, vald, log_alpha, log_alpha_strong,
                              device_strengths=device_strengths)
    return SyntheticPlasma(T, n_e, n_HI, n_HII, n_HeI, h_minus, h2, h2_plus, lt, h_levels=(e_exc, n_lev))
