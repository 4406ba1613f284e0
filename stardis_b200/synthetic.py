"""Synthetic workloads of the named shapes (BASELINE.json configs) -- atmospheres from the two MARCS
models that ship with the reference (structure columns extracted into ``data/atmospheres.npz``), synthetic line
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

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "atmospheres.npz")


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
    # name: (atmosphere, t_scale, lambda range [A], step, n_lines, no_of_thetas)
    "sim10aa": ("sun", 1.0, (6560.0, 6570.0), 0.01, 2000, 20),
    "sim100aa": ("sun", 1.0, (6500.0, 6600.0), 0.01, 2000, 20),
    "solar_full": ("sun", 1.0, (3000.0, 10000.0), 0.01, 300000, 10),
    "astar": ("sun", 9000.0 / 5777.0, (3500.0, 9000.0), 0.01, 100000, 10),
    "coolgiant_ir": ("cool", 1.0, (4000.0, 25000.0), 0.01, 300000, 10),
}


def make_workload(name="solar_full", seed=0, n_lines=None, strong_fraction=0.005, vald=False, lambda_range=None,
                  step=None, log_alpha=(-2.0, 8.0)):
    atm_name, t_scale, lam_rng, dstep, L, n_theta = WORKLOADS[name]
    lam_rng = lambda_range or lam_rng
    step = step or dstep
    L = int(n_lines if n_lines is not None else L)
    atm = load_atmosphere(atm_name, t_scale)
    model = stellar_model_from_atmosphere(atm)
    lam_q, nus = wavelength_grid(lam_rng[0], lam_rng[1], step)
    plasma = create_synthetic_plasma(atm, L, nus.min(), nus.max(), seed=seed, strong_fraction=strong_fraction, vald=vald,
                                     log_alpha=log_alpha)
    return dict(name=name, atmosphere=atm, model=model, plasma=plasma, lambdas=lam_q, nus=nus, no_of_thetas=n_theta)
