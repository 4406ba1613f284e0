"""Broadening (stardis/radiation_field/opacities/opacities_solvers/broadening.py), computed on the device.

Elementwise functions (``calc_doppler_width``, ``calc_n_effective``, ``calc_gamma_linear_stark``,
``calc_gamma_quadratic_stark``, ``calc_gamma_van_der_waals`` and their ``*_cuda`` namesakes, broadening.py:32-547)
broadcast their arguments like the reference's numba ufuncs and run the ``sd_ew_*`` kernels.  ``calculate_broadening``
/ ``calculate_molecule_broadening`` (broadening.py:659-821) take the reference's arguments and run kernel K1
(csrc/k1_broadening.cu) for all (line, depth) pairs at once.
"""
from __future__ import annotations

import logging

import numpy as np

from .... import _lib as L
from .... import units as u
from ....device import default_context

logger = logging.getLogger(__name__)

_FLAG = {"linear_stark": L.LINEAR_STARK, "quadratic_stark": L.QUADRATIC_STARK, "van_der_waals": L.VAN_DER_WAALS,
         "radiation": L.RADIATION}


def broadening_flags(broadening_line_opacity_config):
    """Membership tests of broadening.py:688-691 -> kernel flag word."""
    return sum(bit for name, bit in _FLAG.items() if name in broadening_line_opacity_config)


def calc_doppler_width(nu_line, temperature, atomic_mass, microturbulence=0.0):
    """broadening.py:32-71."""
    return _scalarize(default_context().doppler_width(nu_line, temperature, atomic_mass, float(microturbulence)),
                      nu_line, temperature, atomic_mass)


def calc_n_effective(ion_number, ionization_energy, level_energy):
    """broadening.py:114-146 (NaN where the level lies above the ionisation energy)."""
    return _scalarize(default_context().n_effective(ion_number, ionization_energy, level_energy), ion_number,
                      ionization_energy, level_energy)


def calc_gamma_linear_stark(n_eff_upper, n_eff_lower, electron_density):
    """broadening.py:193-234."""
    return _scalarize(default_context().gamma_linear_stark(n_eff_upper, n_eff_lower, electron_density), n_eff_upper,
                      n_eff_lower, electron_density)


def calc_gamma_quadratic_stark(ion_number, n_eff_upper, n_eff_lower, electron_density, temperature):
    """broadening.py:281-360."""
    return _scalarize(default_context().gamma_quadratic_stark(ion_number, n_eff_upper, n_eff_lower, electron_density,
                                                              temperature), ion_number, n_eff_upper, n_eff_lower,
                      electron_density, temperature)


def calc_gamma_van_der_waals(ion_number, n_eff_upper, n_eff_lower, temperature, h_density):
    """broadening.py:420-490."""
    return _scalarize(default_context().gamma_van_der_waals(ion_number, n_eff_upper, n_eff_lower, temperature, h_density),
                      ion_number, n_eff_upper, n_eff_lower, temperature, h_density)


def _scalarize(out, *args):
    return float(out) if all(np.ndim(a) == 0 for a in args) else out


def _cuda_alias(fn):
    def wrapper(*args, nthreads=256, ret_np_ndarray=True, dtype=float):
        return fn(*args)

    wrapper.__doc__ = f"numba.cuda twin of {fn.__name__} in the reference; same kernel here."
    return wrapper


calc_doppler_width_cuda = _cuda_alias(calc_doppler_width)
calc_n_effective_cuda = _cuda_alias(calc_n_effective)
calc_gamma_linear_stark_cuda = _cuda_alias(calc_gamma_linear_stark)
calc_gamma_quadratic_stark_cuda = _cuda_alias(calc_gamma_quadratic_stark)
calc_gamma_van_der_waals_cuda = _cuda_alias(calc_gamma_van_der_waals)


def _line_columns(lines):
    """Per-line columns from a pandas DataFrame (reference) or a ColumnarLines (fast path)."""
    get = (lambda k: lines[k].values) if hasattr(lines, "columns") else (lambda k: getattr(lines, k))
    has = (lambda k: k in lines.columns) if hasattr(lines, "columns") else (lambda k: getattr(lines, k, None) is not None)
    return get, has


def upload_lines_and_broaden(ctx, lines, alphas_array, masses, stellar_model, stellar_plasma, flags, collective=False):
    """Upload a (sorted, range-selected) line table and run K1.  ``lines``: DataFrame or ColumnarLines.
    ``collective``: this is a multi-GPU run in which every rank is here with the same per-line columns.  "nu": the (L, D)
    strengths are identical on every rank as well (nu sharding): each rank uploads 1/world of the rows and NVLink does the
    rest (``distributed.upload_rows_striped``); "depth": every rank holds its own depth columns of that table (depth
    sharding), which therefore goes up as it is."""
    get, has = _line_columns(lines)
    vald = bool(flags & L.VALD)
    strength = getattr(lines, "strength", None)
    if strength is not None:
        alphas_array = None  # O(L) producer inputs travel instead of the (L, D) table; the device fills it (8f rank 1)
    if collective == "nu" and alphas_array is not None:
        import torch

        from ....distributed import upload_rows_striped

        alphas_array = upload_rows_striped(alphas_array, torch.device("cuda", ctx.device))
        if hasattr(alphas_array, "data_ptr"):
            torch.cuda.current_stream(ctx.device).synchronize()  # the gather ran on torch's stream, K1 runs on ctx's
    cols = dict(nu=get("nu"), mass=masses, atomic_number=np.asarray(get("atomic_number"), dtype=np.int64),
                ion_number=np.asarray(get("ion_number"), dtype=np.int64), ionization_energy=get("ionization_energy"),
                level_energy_upper=get("level_energy_upper"), level_energy_lower=get("level_energy_lower"), A_ul=get("A_ul"))
    if vald and has("stark") and has("waals"):
        cols.update(stark=get("stark"), waals=get("waals"))
    # (The 24 MB of per-line columns go up from every rank directly.  Striping them over the ranks as well
    # (distributed.upload_columns_striped) was measured SLOWER on 2 GPUs -- 30.8 -> 37.8 ms per end-to-end step: the
    # staging copy, the extra collective and its stream synchronisation cost more than the PCIe time they save.)
    ctx.set_lines(cols["nu"], alphas_array, mass=cols["mass"], atomic_number=cols["atomic_number"],
                  ion_number=cols["ion_number"], ionization_energy=cols["ionization_energy"],
                  level_energy_upper=cols["level_energy_upper"], level_energy_lower=cols["level_energy_lower"],
                  A_ul=cols["A_ul"], stark=cols.get("stark"), waals=cols.get("waals"))
    if strength is not None:
        strength.run(ctx)
    ctx.calc_broadening(flags)


def set_device_atmosphere(ctx, stellar_model, stellar_plasma):
    T = u.values_of(stellar_model.temperatures)
    n_e = np.asarray(stellar_plasma.electron_densities.values, dtype=np.float64) if stellar_plasma is not None else None
    n_H = (np.asarray(stellar_plasma.ion_number_density.loc[1, 0].values, dtype=np.float64)
           if stellar_plasma is not None else None)
    vmic = float(u.cgs_values_of(stellar_model.microturbulence))
    ctx.set_atmosphere(T, n_e, n_H, vmic)


def _masses_of(lines, stellar_model):
    get, has = _line_columns(lines)
    if has("mass"):
        return np.asarray(get("mass"), dtype=np.float64)
    return np.asarray(stellar_model.composition.nuclide_masses.loc[get("atomic_number")].values, dtype=np.float64)


def calculate_broadening(lines, stellar_model, stellar_plasma, broadening_line_opacity_config, use_vald_broadening=False):
    """broadening.py:659-732 -> (gammas (L,D), doppler_widths (L,D)) as numpy arrays."""
    flags = broadening_flags(broadening_line_opacity_config) | (L.VALD if use_vald_broadening else 0)
    logger.info("Using VALD broadening parameters." if use_vald_broadening else "Calculating broadening parameters.")
    ctx = default_context()
    ctx.evict()
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    get, _ = _line_columns(lines)
    n = len(get("nu"))
    dummy_alpha = np.zeros((n, stellar_model.no_of_depth_points))
    upload_lines_and_broaden(ctx, lines, dummy_alpha, _masses_of(lines, stellar_model), stellar_model, stellar_plasma, flags)
    return ctx.get(L.BUF_GAMMAS), ctx.get(L.BUF_DOPPLER)


def molecule_masses(lines, stellar_model, stellar_plasma):
    """Sum of the two constituent nuclide masses (broadening.py:808-813)."""
    ions = stellar_plasma.molecule_ion_map.loc[lines.molecule]
    m1 = stellar_model.composition.nuclide_masses.loc[ions.Ion1].values
    m2 = stellar_model.composition.nuclide_masses.loc[ions.Ion2].values
    return np.asarray(m1 + m2, dtype=np.float64)


def molecule_gammas(lines, stellar_model, broadening_line_opacity_config):
    """Non-VALD branch of calculate_molecule_broadening (broadening.py:799-806): gamma = A_ul with shape (L,1) when
    radiation broadening is configured.  Without it the reference dereferences a non-existent attribute
    (``geometry.no_of_depth_points``) and raises AttributeError; the same exception type is raised here."""
    if "radiation" in broadening_line_opacity_config:
        return np.ascontiguousarray(lines.A_ul.values[:, np.newaxis], dtype=np.float64)
    raise AttributeError("'Radial1DGeometry' object has no attribute 'no_of_depth_points'")


def calculate_molecule_broadening(lines, stellar_model, stellar_plasma, broadening_line_opacity_config,
                                  use_vald_broadening=False):
    """broadening.py:735-821 for the branch the reference actually reaches (use_vald_broadening is never passed by
    its caller, opacities_solvers/base.py:469-474).  Doppler widths on the device."""
    if use_vald_broadening:
        raise NotImplementedError("molecular VALD broadening is unreachable in the reference's call graph")
    gammas = molecule_gammas(lines, stellar_model, broadening_line_opacity_config)
    masses = molecule_masses(lines, stellar_model, stellar_plasma)
    T = u.values_of(stellar_model.temperatures)
    vmic = float(u.cgs_values_of(stellar_model.microturbulence))
    dws = default_context().doppler_width(lines.nu.values[:, np.newaxis], T[np.newaxis, :], masses[:, np.newaxis], vmic)
    return gammas, dws


def rotation_broadening(velocity_per_pix, wavelength, flux, v_rot=None, limb_darkening=0.6):
    """Convolve a spectrum with a rotational broadening profile (broadening.py:824-877; only accurate for a constant
    velocity per pixel).  The Gray profile -- 2 (1 - eps) sqrt(1 - (v / v_rot)^2) + (pi / 2) eps (1 - (v / v_rot)^2),
    normalised -- is formed on the host (O(v_rot / velocity_per_pix) numbers); the convolution with the reference's
    ``scipy.ndimage.convolve1d`` semantics (reflecting boundary) runs on the device.  Returns (wavelength, fluxes
    [erg/s/cm2/AA]); a rotational velocity below 1e-5 km/s returns the inputs unchanged."""
    def to_kms(q):
        if not hasattr(q, "to"):
            return float(q)
        try:
            return float(q.to(u.km_s).value)     # stardis_b200.units.Quantity
        except Exception:
            return float(q.to("km/s").value)     # astropy Quantity

    v_pix = to_kms(velocity_per_pix)
    v = 0.0 if v_rot is None else to_kms(v_rot)
    if np.abs(v) < 1e-5:
        return wavelength, flux
    v_rot_by_c = np.maximum(1e-5, np.abs(v)) / (2.99792458e10 * 1e-5)
    half = int(np.round(v / v_pix))
    profile_velocity = np.linspace(-half, half, 2 * half + 1) * v_pix
    profile = np.maximum(0.0, 1.0 - (profile_velocity / v) ** 2)
    rotational = (2 * (1 - limb_darkening) * profile ** 0.5 + 0.5 * np.pi * limb_darkening * profile) / (
        np.pi * v_rot_by_c * (1 - limb_darkening / 3))
    out = default_context().convolve1d_reflect(u.values_of(flux), rotational / rotational.sum())
    return wavelength, u.Quantity(out, "erg/s/cm2/AA")
