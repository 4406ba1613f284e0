"""Synthetic workloads of the named shapes (BASELINE.json configs) -- atmospheres from the two MARCS
models that ship with the reference (structure columns extracted into ``benchdata/atmospheres.npz`` at the repository
root -- fixtures, kept outside the product package), synthetic line
lists and plasma state (``plasma/synthetic.py``).  Used by bench.py and the tests; there is no real atomic data
offline."""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import units as u
from .constants import AMU_CGS, C_CGS
from .io.model.marcs import ATOMIC_WEIGHTS
from .model.base import Composition, Radial1DGeometry, StellarModel
from .plasma.synthetic import create_synthetic_plasma

_FIXTURES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchdata")
_DATA = os.path.join(_FIXTURES, "atmospheres.npz")


def write_cross_section_files(dirpath):
    """The three continuum cross-section tables as text files in the reference's formats (see benchdata/)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_stardis_b200_benchdata", os.path.join(_FIXTURES, "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.write_cross_section_files(dirpath)


def load_atmosphere(name="sun", t_scale=1.0):
    """dict(r, T, pe, pg, density, vmic) deepest -> surface.  ``name`` in {"sun", "cool"}; the hot A-star of
    config #3 has no model file in the reference tree: it is the solar structure with T scaled by ``t_scale``."""
    z = np.load(_DATA)
    return dict(r=-z[f"{name}_depth"], T=z[f"{name}_t"] * t_scale, pe=z[f"{name}_pe"], pg=z[f"{name}_pg"],
                density=z[f"{name}_density"], vmic=float(z[f"{name}_vmic_kms"]) * 1e5, teff=float(z[f"{name}_teff"]) * t_scale)


def stellar_model_from_atmosphere(atm, spherical=False, radius=7.0e10):
    r = atm["r"] + (radius if spherical else 0.0)
    geometry = Radial1DGeometry(u.Quantity(r, u.cm), radius if spherical else None)
    idx = pd.Index(np.arange(1, 93), name="atomic_number")
    comp = Composition(u.Quantity(atm["density"], "g/cm3"), None, pd.Series(ATOMIC_WEIGHTS * AMU_CGS, index=idx))
    return StellarModel(u.Quantity(atm["T"], u.K), geometry, comp, spherical=spherical,
                        microturbulence=u.Quantity(atm["vmic"] / 1e5, u.km_s))


def wavelength_grid(lambda_min, lambda_max, step=0.01):
    """np.arange(lambda_min, lambda_max, step) Angstrom -> descending frequencies (run_stardis, base.py:34)."""
    lam = np.arange(lambda_min, lambda_max, step)
    return u.Quantity(lam, u.AA), C_CGS / (lam * 1e-8)


# BASELINE.json configs -> shapes
WORKLOADS = {
    # name: (atmosphere, t_scale, lambda range [A], step, n_lines, no_of_thetas[, line-list overrides])
    "sim10aa": ("sun", 1.0, (6560.0, 6570.0), 0.01, 2000, 20),
    "sim100aa": ("sun", 1.0, (6500.0, 6600.0), 0.01, 2000, 20),
    "solar_full": ("sun", 1.0, (3000.0, 10000.0), 0.01, 300000, 10),
    # the flagship grid with a WEAK-line list: windows of 20..~6000 pixels, no whole-grid wings -- the regime in which
    # the line cores (Humlicek regions II-IV), not the far field, set the rate
    "solar_weak": ("sun", 1.0, (3000.0, 10000.0), 0.01, 300000, 10, dict(log_alpha=(-6.0, 2.0), strong_fraction=0.0)),
    "astar": ("sun", 9000.0 / 5777.0, (3500.0, 9000.0), 0.01, 100000, 10),
    "coolgiant_ir": ("cool", 1.0, (4000.0, 25000.0), 0.01, 300000, 10),
}


def make_workload(name="solar_full", seed=0, n_lines=None, strong_fraction=0.005, vald=False, lambda_range=None,
                  step=None, log_alpha=(-2.0, 8.0), device_strengths=False):
    atm_name, t_scale, lam_rng, dstep, L, n_theta = WORKLOADS[name][:6]
    if len(WORKLOADS[name]) > 6:
        ov = WORKLOADS[name][6]
        log_alpha = ov.get("log_alpha", log_alpha)
        strong_fraction = ov.get("strong_fraction", strong_fraction)
    lam_rng = lambda_range or lam_rng
    step = step or dstep
    L = int(n_lines if n_lines is not None else L)
    atm = load_atmosphere(atm_name, t_scale)
    model = stellar_model_from_atmosphere(atm)
    lam_q, nus = wavelength_grid(lam_rng[0], lam_rng[1], step)
    plasma = create_synthetic_plasma(atm, L, nus.min(), nus.max(), seed=seed, strong_fraction=strong_fraction, vald=vald,
                                     log_alpha=log_alpha, device_strengths=device_strengths)
    return dict(name=name, atmosphere=atm, model=model, plasma=plasma, lambdas=lam_q, nus=nus, no_of_thetas=n_theta)


def sweep_models(w, n_models=64):
    """BASELINE.json configs[4] (grid sweep): ``n_models`` variations of a workload's atmosphere on a Teff x log g x
    [Fe/H] grid (4 x 4 x 4 for 64) -- temperatures scaled by Teff / Teff_0, pressures by 10^(dlogg / 2), line strengths
    by 10^[Fe/H] with the Boltzmann depth trend re-evaluated for the scaled temperatures.  The line list (positions,
    levels) is shared.  Returns ``model(m)`` -> the per-model inputs of the hot path: T, n_e, n_H (D,), alpha_line (L, D),
    continuum descriptors (bf_prefix, ff_coef, electron, tables) and a description."""
    from .constants import KB_CGS
    from .io.config import Configuration
    from .radiation_field.opacities.opacities_solvers import base as ob

    side = max(1, round(n_models ** (1.0 / 3.0)))
    teffs = np.linspace(5000.0, 6500.0, side)
    dloggs = np.linspace(-0.5, 1.0, side)
    fehs = np.linspace(-1.0, 0.3, side)
    atm0, lt = w["atmosphere"], w["plasma"]._line_table
    a0 = lt.alpha_line[:, int(np.argmax(atm0["T"]))].copy()  # the Boltzmann factor is normalised at the hottest depth
    hm_file = {"Hminus_bf": write_cross_section_files(os.path.join(_tmpdir(), "tables"))["Hminus_bf"]}
    species = Configuration({"H_I": {}})

    def model(m):
        i, j, k = (m // (side * side)) % side, (m // side) % side, m % side
        atm = dict(atm0)
        atm["T"] = atm0["T"] * (teffs[i] / atm0["teff"])
        atm["pe"] = atm0["pe"] * 10.0 ** (0.5 * dloggs[j])
        atm["pg"] = atm0["pg"] * 10.0 ** (0.5 * dloggs[j])
        T = atm["T"]
        alpha = (a0 * 10.0 ** fehs[k])[:, None] * np.exp(-lt.level_energy_lower[:, None] / KB_CGS * (1.0 / T[None, :] - 1.0 / T.max()))
        sm = stellar_model_from_atmosphere(atm)
        pl = create_synthetic_plasma(atm, 0, w["nus"].min(), w["nus"].max(), seed=0)
        tables, _ = ob.file_tables(pl, sm, hm_file)
        _, bf_prefix = ob.bf_descriptor(pl, species)
        return dict(T=T, n_e=pl.electron_densities.values.copy(), n_H=pl.ion_number_density.loc[1, 0].values.copy(),
                    alpha_line=np.ascontiguousarray(alpha), continuum=(bf_prefix, ob.ff_descriptor(pl, sm, species),
                                                                       ob.electron_descriptor(pl), tables),
                    desc=f"Teff {teffs[i]:.0f} K, dlogg {dloggs[j]:+.1f}, [Fe/H] {fehs[k]:+.1f}")

    return model


def _tmpdir():
    import tempfile

    return tempfile.mkdtemp(prefix="sdb200_sweep_")
