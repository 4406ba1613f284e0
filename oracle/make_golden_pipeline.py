"""Pipeline-level goldens: the reference's calc_alphas (incl. its pandas line selection, broadening, Voigt loop and
every continuum term) and raytrace, run UNMODIFIED on the duck-typed synthetic plasma/model of the product package.

TEST INFRASTRUCTURE (build container only; see make_golden.py).  Writes
  tests/golden/pipeline_golden.npz   reference outputs per configuration + fingerprints of the seeded inputs
"""
from __future__ import annotations

import os
import types

import numpy as np

from oracle.ref_shim import FakeQuantity, REFERENCE_ROOT
from oracle import oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
DATA = os.path.join(REFERENCE_ROOT, "stardis", "data")
TABLE_FILES = {"Hminus_bf": "h_minus_bf_W1979.dat", "Hminus_ff": "h_minus_ff_B1987.dat", "H2plus_bf": "h2_plus_bf_S1994.dat"}

DEPTH_ROWS = [0, 11, 27, 40, 50, 55]

# configurations exercised (opacity section of the YAML, as dicts)
CASES = {
    # benchmark_config.yml-like: H- bf file, H I bf+ff, all four broadenings
    # (use_vald_broadening: false makes the reference drop auto-ionising lines, opacities_solvers/base.py:413-421)
    "bench": dict(file={"Hminus_bf": "Hminus_bf"}, bf={"H_I": {}}, ff={"H_I": {}}, rayleigh=[],
                  disable_electron_scattering=False,
                  line=dict(disable=False, broadening=["radiation", "linear_stark", "quadratic_stark", "van_der_waals"],
                            vald_linelist=dict(use_vald_broadening=False))),
    # stardis_test_config_broadening.yml-like: three file opacities + Rayleigh
    "broadening": dict(file={"Hminus_bf": "Hminus_bf", "Hminus_ff": "Hminus_ff", "H2plus_bf": "H2plus_bf"}, bf={"H_I": {}},
                       ff={"H_I": {}}, rayleigh=["H", "He", "H2"], disable_electron_scattering=False,
                       line=dict(disable=False, broadening=["radiation", "linear_stark", "quadratic_stark", "van_der_waals"])),
    # stardis_test_config.yml-like: no electron scattering, no broadening
    "plain": dict(file={}, bf={"H_I": {}}, ff={"H_I": {}}, rayleigh=[], disable_electron_scattering=True,
                  line=dict(disable=False, broadening=[], vald_linelist=dict(use_vald_broadening=False))),
    # VALD line list with VALD broadening parameters
    "vald": dict(file={}, bf={}, ff={}, rayleigh=["H"], disable_electron_scattering=False,
                 line=dict(disable=False, broadening=["radiation", "quadratic_stark", "van_der_waals"],
                           vald_linelist=dict(use_linelist=True, use_vald_broadening=True))),
    # lines disabled
    "nolines": dict(file={"Hminus_bf": "Hminus_bf"}, bf={"H_I": {}}, ff={}, rayleigh=[], disable_electron_scattering=False,
                    line=dict(disable=True)),
}


def case_inputs(name, opacity, table_paths):
    """Seeded inputs of a case: shared by this generator and tests/test_gpu_pipeline.py."""
    from stardis_b200.io.config import Configuration, validate_config
    from stardis_b200.synthetic import load_atmosphere, stellar_model_from_atmosphere, wavelength_grid
    from stardis_b200.plasma.synthetic import create_synthetic_plasma

    op = dict(opacity)
    op["file"] = {k: table_paths[v] for k, v in opacity["file"].items()}
    cfg = Configuration(validate_config(dict(stardis_config_version=1.0, atom_data="synthetic:300",
                                             input_model=dict(type="marcs", fname="unused"), opacity=op,
                                             no_of_thetas=5, result_options=dict(return_radiation_field=True))))
    atm = load_atmosphere("sun")
    spherical = name == "broadening"
    model = stellar_model_from_atmosphere(atm, spherical=spherical)
    lam, nus = wavelength_grid(6556.0, 6570.0, 0.02)
    vald = cfg.opacity.line.vald_linelist.use_linelist
    plasma = create_synthetic_plasma(atm, 300, nus.min() * 0.999, nus.max() * 1.001, seed=11, strong_fraction=0.02, vald=vald,
                                     log_alpha=(-2.0, 6.0), log_alpha_strong=(7.0, 9.0))
    # A few auto-ionising lines where the reference drops them (use_vald_broadening false) or where their broadening
    # does not involve n_eff (VALD parameters).  With the default flag they would carry NaN gammas, and what the
    # reference's fastmath-compiled line loop does with NaN depends on the numba specialisation (observed: NaN
    # written to +-10 pixels for one array layout, nothing written for another) -- not a parity target.
    lt = plasma._line_table
    if name in ("bench", "plain", "vald"):
        lt.level_energy_upper[::47] = lt.ionization_energy[::47] * 1.02
        if vald:
            lt.waals[::47] = -7.5  # scaled-gamma form: no n_eff involved
    return cfg, model, plasma, nus


def main(R):
    os.makedirs(OUT, exist_ok=True)
    table_paths = {src: os.path.join(DATA, fn) for src, fn in TABLE_FILES.items()}

    out = {}
    for name, opacity in CASES.items():
        cfg, model, plasma, nus = case_inputs(name, opacity, table_paths)
        # the reference reads astropy-like quantities: wrap the model for it
        ref_model = types.SimpleNamespace(
            temperatures=FakeQuantity(np.asarray(model.temperatures), "K"), no_of_depth_points=model.no_of_depth_points,
            geometry=types.SimpleNamespace(r=np.asarray(model.geometry.r), dist_to_next_depth_point=model.geometry.dist_to_next_depth_point,
                                           reference_r=model.geometry.reference_r),
            spherical=model.spherical, composition=model.composition,
            microturbulence=types.SimpleNamespace(cgs=types.SimpleNamespace(value=float(model.microturbulence.cgs.value))))
        freqs = FakeQuantity(nus, "Hz")
        x, w = np.polynomial.legendre.leggauss(cfg.no_of_thetas)
        srf = types.SimpleNamespace(frequencies=freqs, opacities=R.opacities_container.Opacities(freqs, ref_model),
                                    thetas=(x / 2) + 0.5 * np.pi / 2, I_nus_weights=w * np.pi / 2,
                                    source_function=R.blackbody.blackbody_flux_at_nu, track_individual_intensities=True,
                                    F_nu=np.zeros((model.no_of_depth_points, len(nus))),
                                    I_nus=np.zeros((model.no_of_depth_points, len(nus), cfg.no_of_thetas)))
        total = R.opac.calc_alphas(plasma, ref_model, srf, cfg.opacity)
        R.solver.raytrace(ref_model, srf)
        for key, val in srf.opacities.opacities_dict.items():
            val = np.asarray(val, dtype=np.float64)
            # per-term (D, N) arrays are kept at a few depth rows only (fixture size); total and F_nu in full
            out[f"{name}__{key}"] = val[DEPTH_ROWS] if val.ndim == 2 and val.shape[1] == len(nus) else val
        out[f"{name}__total"] = np.asarray(total)
        out[f"{name}__F_nu"] = srf.F_nu
        out[f"{name}__I_nus_emergent"] = srf.I_nus[-1]
        out[f"{name}__fingerprint"] = np.array([plasma._line_table.nu.sum(), plasma._line_table.alpha_line.sum(),
                                                plasma.electron_densities.values.sum(), nus.sum()])
        print(name, {k: np.shape(v) for k, v in srf.opacities.opacities_dict.items()})
    np.savez_compressed(os.path.join(OUT, "pipeline_golden.npz"), **out)


if __name__ == "__main__":
    from oracle.ref_shim import load_reference

    main(load_reference())
