"""Opacity driver: B200 implementation behind the reference's function signatures.

Mirror of stardis/radiation_field/opacities/opacities_solvers/base.py (``calc_alphas`` :630-740 and the functions
it calls).  The host side only assembles O(D) + O(levels) coefficient vectors and selects the line window; every
(depth, nu) and (line, depth) quantity is produced by the kernels of libstardis_b200.so:

    K1  sd_calc_broadening   gammas / Doppler widths per (line, depth)            csrc/k1_broadening.cu
    K2  sd_calc_alpha_line   windowed Voigt accumulation -> alpha_line[d, nu]     csrc/k2_lines.cu
    K3  sd_calc_continuum    file/bf/ff/Rayleigh/electron terms + total, fused    csrc/k3_continuum.cu

There is no CPU fallback: without the library or a GPU these functions raise.
"""
from __future__ import annotations

import logging
import weakref

import numpy as np

from .... import _lib as L
from .... import units as u
from ....constants import BF_CONSTANT, FF_CONSTANT, H_CGS, RYDBERG_FREQUENCY, SIGMA_T_CGS
from ....device import default_context
from ....device_array import DeviceArray
from ....plasma.columnar import ColumnarLines
from .broadening import (broadening_flags, molecule_gammas, molecule_masses, set_device_atmosphere,
                         upload_lines_and_broaden)
from .util import get_number_density, table_descriptor

logger = logging.getLogger(__name__)

VACUUM_ELECTRIC_PERMITTIVITY = 1 / (4 * np.pi)
RAYLEIGH_UPPER_BOUND_HZ = 2.3e15


# ------------------------------------------------------------------------------------------ helpers
def _ctx_of(stellar_radiation_field=None):
    ctx = getattr(stellar_radiation_field, "device_context", None)
    return ctx if ctx is not None else default_context()


def _shard_of(stellar_radiation_field, n):
    shard = getattr(stellar_radiation_field, "shard", None)
    return (0, n) if shard is None else (int(shard[0]), int(shard[1]))


def _depth_values(x):
    return np.asarray(getattr(x, "values", x), dtype=np.float64)


def _zeros_lazy(shape):
    return DeviceArray(None, None, shape, fetch=lambda: np.zeros(shape))


# ------------------------------------------------------------------------------------------ descriptors (host, O(D))
def file_tables(stellar_plasma, stellar_model, opacity_file_config):
    """One device table per requested file opacity (calc_alpha_file, base.py:40-70; config order preserved)."""
    T = u.values_of(stellar_model.temperatures)
    tables, names = [], []
    for opacity_source, fpath in opacity_file_config.items():
        number_density, _, _ = get_number_density(stellar_plasma, opacity_source)
        tables.append(table_descriptor(fpath, opacity_source, T, _depth_values(number_density)))
        names.append(f"alpha_file_{opacity_source}")
    return tables, names


def bf_descriptor(stellar_plasma, species):
    """Hydrogenic bound-free edges (calc_alpha_bf base.py:178-239, calc_contribution_bf :242-271).

    Every level of every configured species contributes BF * (ion+1)^4 * n_level[d] / n^5 above its cutoff
    frequency (E_ion - E_exc) / h, n^5 = ((ion+1) sqrt(nu_Ryd / nu_cut))^5.  Returned sorted by cutoff with depth-wise
    prefix sums so that the kernel needs one binary search per frequency."""
    cut, rows = [], []
    D = None
    for spec in species.keys():
        _, atomic_number, ion_number = get_number_density(stellar_plasma, spec + "_bf")
        ionization_energy = float(stellar_plasma.ionization_data.loc[(atomic_number, ion_number + 1)])
        lv = stellar_plasma.levels
        sel = (lv.get_level_values(0) == atomic_number) & (lv.get_level_values(1) == ion_number)
        exc = np.asarray(stellar_plasma.excitation_energy.loc[lv[sel]].values, dtype=np.float64)
        dens = np.asarray(stellar_plasma.level_number_density.loc[lv[sel]].values, dtype=np.float64)
        D = dens.shape[1]
        cutoff = (ionization_energy - exc) / H_CGS
        n5 = ((ion_number + 1) * np.sqrt(RYDBERG_FREQUENCY / cutoff)) ** 5
        cut.append(cutoff)
        rows.append(BF_CONSTANT * (ion_number + 1) ** 4 * dens / n5[:, None])
    if not cut:
        return None, None
    cut = np.concatenate(cut)
    rows = np.vstack(rows)
    order = np.argsort(cut, kind="stable")
    prefix = np.zeros((len(cut) + 1, D))
    np.cumsum(rows[order], axis=0, out=prefix[1:])
    return np.ascontiguousarray(cut[order]), np.ascontiguousarray(prefix)


def ff_descriptor(stellar_plasma, stellar_model, species):
    """calc_alpha_ff (base.py:274-317): coef[d] = sum_species (n_e n_ion / sqrt(T)) * (FF Z^2)."""
    if not len(species):
        return None
    T = u.values_of(stellar_model.temperatures)
    coef = np.zeros(T.shape[0])
    for spec in species.keys():
        number_density, _, ion_number = get_number_density(stellar_plasma, spec + "_ff")
        coef += (_depth_values(number_density) / np.sqrt(T)) * (FF_CONSTANT * ion_number**2)
    return coef


def rayleigh_descriptor(stellar_plasma, stellar_model, species):
    """Coefficients of calc_alpha_rayleigh (base.py:101-125)."""
    D = stellar_model.no_of_depth_points
    c4, c6, c8 = np.zeros(D), np.zeros(D), np.zeros(D)
    if "H" in species:
        n = np.array(stellar_plasma.ion_number_density.loc[1, 0])
        c4 += 20.24 * n; c6 += 239.2 * n; c8 += 2256 * n
    if "He" in species:
        n = np.array(stellar_plasma.ion_number_density.loc[2, 0])
        c4 += 1.913 * n; c6 += 4.52 * n; c8 += 7.90 * n
    if "H2" in species:
        n = np.array(stellar_plasma.h2_density)
        c4 += 28.39 * n; c6 += 215.0 * n; c8 += 1303 * n
    return c4, c6, c8


def electron_descriptor(stellar_plasma):
    return SIGMA_T_CGS * np.asarray(stellar_plasma.electron_densities.values, dtype=np.float64)


def clip_rayleigh_frequencies(tracing_nus):
    """base.py:98-99 zeroes frequencies above 2.3e15 Hz IN PLACE in the caller's array (a quirk that is kept:
    everything evaluated afterwards sees the modified grid).  Returns True if anything changed."""
    vals = u.values_of(tracing_nus)
    mask = vals > RAYLEIGH_UPPER_BOUND_HZ
    if mask.any():
        vals[mask] = 0
        return True
    return False


# ------------------------------------------------------------------------------------------ single-term entry points
def _single_term(stellar_plasma, stellar_model, tracing_nus, source_index, **desc):
    ctx = default_context()
    ctx.evict()
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    ctx.set_grid(u.values_of(tracing_nus))
    ctx.calc_continuum(store_mask=1 << source_index, **desc)
    return ctx.get(L.BUF_SOURCE0 + source_index)


def calc_alpha_file(stellar_plasma, stellar_model, tracing_nus, opacity_source, fpath):
    """base.py:40-70 -> (D, N)."""
    tables, _ = file_tables(stellar_plasma, stellar_model, {opacity_source: fpath})
    return _single_term(stellar_plasma, stellar_model, tracing_nus, L.SRC_TABLE0, tables=tables)


def calc_alpha_rayleigh(stellar_plasma, stellar_model, tracing_nus, species):
    """base.py:74-135 -> (D, N); mutates ``tracing_nus`` in place like the reference."""
    out = _single_term(stellar_plasma, stellar_model, tracing_nus, L.SRC_RAYLEIGH,
                       rayleigh=rayleigh_descriptor(stellar_plasma, stellar_model, species))
    clip_rayleigh_frequencies(tracing_nus)
    return out


def calc_alpha_electron(stellar_plasma, stellar_model, tracing_nus, disable_electron_scattering=False):
    """base.py:139-174 -> (D, N), or the integer 0 when disabled."""
    if disable_electron_scattering:
        return 0
    return _single_term(stellar_plasma, stellar_model, tracing_nus, L.SRC_ELECTRON, electron=electron_descriptor(stellar_plasma))


def calc_alpha_bf(stellar_plasma, stellar_model, tracing_nus, species):
    """base.py:178-239 -> (D, N)."""
    cut, prefix = bf_descriptor(stellar_plasma, species)
    if cut is None:
        return np.zeros((stellar_model.no_of_depth_points, len(tracing_nus)))
    return _single_term(stellar_plasma, stellar_model, tracing_nus, L.SRC_BF, bf_nu_cut=cut, bf_prefix=prefix)


def calc_alpha_ff(stellar_plasma, stellar_model, tracing_nus, species):
    """base.py:274-317 -> (D, N)."""
    coef = ff_descriptor(stellar_plasma, stellar_model, species)
    if coef is None:
        return np.zeros((stellar_model.no_of_depth_points, len(tracing_nus)))
    return _single_term(stellar_plasma, stellar_model, tracing_nus, L.SRC_FF, ff_coef=coef)


def gaunt_times_departure(tracing_nus, temperatures, gaunt_fpath, departure_fpath):
    """To be implemented (base.py:320-324: a stub in the reference as well)."""
    pass


# ------------------------------------------------------------------------------------------ line opacity
_columnar_cache = weakref.WeakKeyDictionary()


def columnar_lines_of(stellar_plasma, use_vald):
    """The plasma's nu-sorted columnar line table: taken from ``stellar_plasma.line_table`` when the plasma carries
    one, otherwise assembled ONCE from the pandas tables the reference reads (base.py:362-407) and cached."""
    native = getattr(stellar_plasma, "line_table", None)
    if native is not None:
        return native
    try:
        per_plasma = _columnar_cache.setdefault(stellar_plasma, {})
    except TypeError:  # plasma object is not weak-referenceable
        per_plasma = stellar_plasma.__dict__.setdefault("_stardis_b200_columnar", {})
    if use_vald not in per_plasma:
        per_plasma[use_vald] = ColumnarLines.from_plasma(stellar_plasma, use_vald)
    return per_plasma[use_vald]


def select_lines(stellar_plasma, stellar_model, tracing_nus, line_opacity_config, extent=None):
    """Lines inside [min nu, max nu] (base.py:392-407), auto-ionising ones dropped only when VALD broadening is
    off (:413-421), masses attached (broadening.py:723-730).  ``extent``: (min, max) of the grid when the caller has
    them already (two passes over the grid otherwise)."""
    use_vald = line_opacity_config.vald_linelist.use_linelist
    table = columnar_lines_of(stellar_plasma, use_vald)
    if not line_opacity_config.vald_linelist.use_vald_broadening:
        table = table.without_autoionizing()
    table.with_masses(stellar_model.composition.nuclide_masses)
    if extent is None:
        nus = u.values_of(tracing_nus)
        extent = (nus.min(), nus.max())
    return table.in_range(extent[0], extent[1])


def _line_flags(line_opacity_config):
    use_vald_broadening = (line_opacity_config.vald_linelist.use_vald_broadening
                           and line_opacity_config.vald_linelist.use_linelist)  # base.py:428-429
    return broadening_flags(line_opacity_config.broadening) | (L.VALD if use_vald_broadening else 0)


def _device_line_opacity(ctx, stellar_plasma, stellar_model, tracing_nus, line_opacity_config, collective=False, extent=None):
    """K1 + K2 for the atomic lines on an already prepared context (atmosphere + grid set).  Returns the number of
    lines used."""
    lines = select_lines(stellar_plasma, stellar_model, tracing_nus, line_opacity_config, extent=extent)
    upload_lines_and_broaden(ctx, lines, lines.alpha_line, lines.mass, stellar_model, stellar_plasma,
                             _line_flags(line_opacity_config), collective=collective)
    logger.info("Calculating line opacities at spectral points.")
    ctx.calc_alpha_line(0)
    return len(lines)


def calc_alpha_line_at_nu(stellar_plasma, stellar_model, tracing_nus, line_opacity_config):
    """base.py:328-441 -> (alpha_line_at_nu (D,N), gammas (L,D), doppler_widths (L,D)) or (0, 0, 0)."""
    if line_opacity_config.disable:
        return 0, 0, 0
    ctx = default_context()
    ctx.evict()
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    ctx.set_grid(u.values_of(tracing_nus))
    n_lines = _device_line_opacity(ctx, stellar_plasma, stellar_model, tracing_nus, line_opacity_config)
    D = stellar_model.no_of_depth_points
    alpha = ctx.get(L.BUF_ALPHA_LINE)
    if n_lines == 0:
        return alpha, np.zeros((0, D)), np.zeros((0, D))
    return alpha, ctx.get(L.BUF_GAMMAS), ctx.get(L.BUF_DOPPLER)


def _molecular_inputs(stellar_plasma, stellar_model, tracing_nus, line_opacity_config):
    """Selection and broadening inputs of calc_molecular_alpha_line_at_nu (base.py:444-484)."""
    nus = u.values_of(tracing_nus)
    lines = stellar_plasma.molecule_lines_from_linelist
    lines_sorted = lines.sort_values("nu")
    in_range = lines_sorted[lines_sorted.nu.between(nus.min(), nus.max())]
    alphas_and_nu = stellar_plasma.molecule_alpha_line_from_linelist.sort_values("nu")
    alphas_array = alphas_and_nu[alphas_and_nu.nu.between(nus.min(), nus.max())].drop(labels="nu", axis=1).to_numpy()
    gammas = molecule_gammas(in_range, stellar_model, line_opacity_config.broadening)
    masses = molecule_masses(in_range, stellar_model, stellar_plasma)
    return in_range, np.ascontiguousarray(alphas_array, dtype=np.float64), gammas, masses


def _device_molecular_opacity(ctx, stellar_plasma, stellar_model, tracing_nus, line_opacity_config):
    lines, alphas_array, gammas, masses = _molecular_inputs(stellar_plasma, stellar_model, tracing_nus, line_opacity_config)
    T = u.values_of(stellar_model.temperatures)
    vmic = float(u.cgs_values_of(stellar_model.microturbulence))
    line_nus = np.ascontiguousarray(lines.nu.to_numpy(), dtype=np.float64)
    dws = ctx.doppler_width(line_nus[:, None], T[None, :], masses[:, None], vmic)
    ctx.set_lines(line_nus, alphas_array, mass=masses)
    ctx.set_broadening(gammas, dws)
    ctx.calc_alpha_line(1)
    return gammas, dws


def calc_molecular_alpha_line_at_nu(stellar_plasma, stellar_model, tracing_nus, line_opacity_config):
    """base.py:444-484."""
    if line_opacity_config.disable:
        return 0, 0, 0
    ctx = default_context()
    ctx.evict()
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    ctx.set_grid(u.values_of(tracing_nus))
    gammas, dws = _device_molecular_opacity(ctx, stellar_plasma, stellar_model, tracing_nus, line_opacity_config)
    return ctx.get(L.BUF_ALPHA_MOLECULE), gammas, dws


def calc_alan_entries(no_of_depth_points, tracing_nus_values, line_nus, doppler_widths, gammas, alphas_array):
    """base.py:487-592 with the reference's argument list; arrays in, (D, N) array out, computed by K2.

    ``tracing_nus_values`` must be descending and ``line_nus`` ascending, as in the reference.  A zero Doppler width
    inside a non-empty window raises ZeroDivisionError (voigt.py:148 does, inside the numba kernel)."""
    ctx = default_context()
    ctx.evict()
    D = int(no_of_depth_points)
    ctx.set_atmosphere(np.ones(D))
    ctx.set_grid(np.asarray(tracing_nus_values, dtype=np.float64))
    ctx.set_lines(np.asarray(line_nus, dtype=np.float64), np.asarray(alphas_array, dtype=np.float64))
    ctx.set_broadening(np.asarray(gammas, dtype=np.float64), np.asarray(doppler_widths, dtype=np.float64))
    ctx.calc_alpha_line(0)
    if len(line_nus) and ctx.line_stats()["zero_doppler_pairs"]:
        raise ZeroDivisionError("division by zero")
    return ctx.get(L.BUF_ALPHA_LINE)


def _calc_alan_entries(delta_nus, doppler_widths_at_depth_point, gammas_at_depth_point, alphas_at_depth_point):
    """base.py:595-627: phi * alpha for one (line, depth) over a window of frequency offsets."""
    from .voigt import voigt_profile

    return voigt_profile(delta_nus, doppler_widths_at_depth_point, gammas_at_depth_point) * alphas_at_depth_point


# ------------------------------------------------------------------------------------------ driver
def opacity_context_of(ctx):
    """Second context on the same device for the depth-sharded opacity stages of a multi-GPU run (the formal solution
    keeps ``ctx``, which holds all depth points of this rank's pixel range)."""
    from ....device import DeviceContext

    if getattr(ctx, "_opacity_ctx", None) is None:
        ctx._opacity_ctx = DeviceContext(ctx.device)
    return ctx._opacity_ctx


def calc_alphas(stellar_plasma, stellar_model, stellar_radiation_field, opacity_config, store_components=True):
    """Calculates every opacity term, stores them in the radiation field and returns the total (base.py:630-740).

    One pass on the device: K1 -> K2 (atomic, then molecular lines) -> K3 (all continuum terms + total).  The
    entries of ``opacities_dict`` keep the reference's keys and order; they are ``DeviceArray`` objects that turn
    into numpy arrays when touched.  With ``store_components=False`` the per-term arrays are not kept in HBM (only the
    total is); touching one then recomputes that single term.

    Multi-GPU: ``stellar_radiation_field.shard`` = (p0, p1) alone evaluates that pixel range of the global grid (nu
    sharding, no exchange); with ``depth_shard`` = (rank, world) as well, the opacity stages run for the depth points
    rank, rank + world, ... on the WHOLE grid and one all-to-all redistributes the result to the pixel ranges (see
    ``stardis_b200.distributed``)."""
    srf = stellar_radiation_field
    depth_shard = getattr(srf, "depth_shard", None)
    if depth_shard is not None and int(depth_shard[1]) > 1:
        return _calc_alphas_depth_sharded(stellar_plasma, stellar_model, srf, opacity_config, store_components)
    ctx = _ctx_of(srf)
    ctx.evict()
    nus_q = srf.frequencies
    nus = u.values_of(nus_q)
    N = nus.shape[0]
    D = stellar_model.no_of_depth_points
    p0, p1 = _shard_of(srf, N)
    W = p1 - p0

    extent = (float(nus.min()), float(nus.max()))
    if not extent[1] <= RAYLEIGH_UPPER_BOUND_HZ:
        raise NotImplementedError(
            "frequency grids reaching above 2.3e15 Hz (lambda < 1303 A) trigger the reference's in-place frequency "
            "zeroing (base.py:98-99), after which its own line and formal-solver steps are ill-defined; not supported")
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    ctx.set_grid(nus, p0, p1)
    n_lines, mol = _device_opacity_pass(ctx, stellar_plasma, stellar_model, nus_q, opacity_config, store_components,
                                        collective="nu" if getattr(srf, "shard", None) is not None else False, extent=extent)
    ctx.owner = getattr(srf, "token", None)

    def shard_array(which):
        return ctx.track(DeviceArray(ctx, which, (D, W)))

    def line_table_array(which):
        return ctx.track(DeviceArray(ctx, which, (n_lines, D)))

    _publish(srf, stellar_plasma, stellar_model, nus_q, opacity_config, store_components, n_lines, mol, (D, W), (p0, p1),
             shard_array, line_table_array)
    # ---- total (Opacities.calc_total_alphas accumulates into the existing array, opacities/base.py:24-28)
    previous = srf.opacities._total
    total = shard_array(L.BUF_TOTAL)
    if previous is not None and np.any(np.asarray(previous) != 0):
        total = np.asarray(previous) + total.numpy()
        ctx.set_total(total)
    srf.opacities.total_alphas = total
    return total


def _device_opacity_pass(ctx, stellar_plasma, stellar_model, nus_q, opacity_config, store_components, collective=False,
                         extent=None):
    """K1 + K2 (atomic, molecular) + K3 on a prepared context (atmosphere and grid set).  Continuum descriptors are
    assembled on the host (O(D)); evaluation order follows base.py:655-736.  Returns (n_lines or None, molecular
    (gammas, doppler_widths) or None)."""
    line_cfg = opacity_config.line
    n_lines, mol = None, None
    if not line_cfg.disable:
        # nu-sharded run (every rank of the job is here with the same line table): stripe the big upload
        n_lines = _device_line_opacity(ctx, stellar_plasma, stellar_model, nus_q, line_cfg, collective=collective, extent=extent)
        if line_cfg.include_molecules:
            mol = _device_molecular_opacity(ctx, stellar_plasma, stellar_model, nus_q, line_cfg)
    # the line kernels are running now: the host assembles the O(D) continuum descriptors (pandas) in their shadow
    tables, _ = file_tables(stellar_plasma, stellar_model, opacity_config.file)
    bf_cut, bf_prefix = bf_descriptor(stellar_plasma, opacity_config.bf)
    ff_coef = ff_descriptor(stellar_plasma, stellar_model, opacity_config.ff)
    rayleigh = rayleigh_descriptor(stellar_plasma, stellar_model, opacity_config.rayleigh)
    electron = None if opacity_config.disable_electron_scattering else electron_descriptor(stellar_plasma)
    ctx.calc_continuum(bf_nu_cut=bf_cut, bf_prefix=bf_prefix, ff_coef=ff_coef, rayleigh=rayleigh, electron=electron,
                       tables=tables, store_mask=0xFFFF if store_components else 0)
    return n_lines, mol


def _publish(srf, stellar_plasma, stellar_model, nus_q, opacity_config, store_components, n_lines, mol, shape, pixel_range,
             shard_array, line_table_array):
    """Fill ``opacities_dict`` with the reference's keys in the reference's order (base.py:655-736).  ``shard_array(which)``
    / ``line_table_array(which)`` wrap a stored device result; terms that were not stored are recomputed when touched."""
    od = srf.opacities.opacities_dict
    D, W = shape
    p0, p1 = pixel_range
    line_cfg = opacity_config.line
    have_bf = len(opacity_config.bf) > 0
    have_ff = len(opacity_config.ff) > 0

    def term(source_index, recompute):
        if store_components:
            return shard_array(L.BUF_SOURCE0 + source_index)
        return DeviceArray(None, None, (D, W), fetch=lambda: recompute()[:, p0:p1])

    for k, (src, fpath) in enumerate(list(opacity_config.file.items())):
        od[f"alpha_file_{src}"] = term(L.SRC_TABLE0 + k, lambda src=src, fpath=fpath: calc_alpha_file(
            stellar_plasma, stellar_model, nus_q, src, fpath))
    od["alpha_bf"] = (term(L.SRC_BF, lambda: calc_alpha_bf(stellar_plasma, stellar_model, nus_q, opacity_config.bf))
                      if have_bf else _zeros_lazy((D, W)))
    od["alpha_ff"] = (term(L.SRC_FF, lambda: calc_alpha_ff(stellar_plasma, stellar_model, nus_q, opacity_config.ff))
                      if have_ff else _zeros_lazy((D, W)))
    od["alpha_rayleigh"] = term(L.SRC_RAYLEIGH, lambda: calc_alpha_rayleigh(stellar_plasma, stellar_model, nus_q, opacity_config.rayleigh))
    od["alpha_electron"] = (0 if opacity_config.disable_electron_scattering else
                            term(L.SRC_ELECTRON, lambda: calc_alpha_electron(stellar_plasma, stellar_model, nus_q)))
    if line_cfg.disable:
        od["alpha_line_at_nu"] = 0
        od["alpha_line_at_nu_gammas"] = 0
        od["alpha_line_at_nu_doppler_widths"] = 0
        return

    def _atomic(which):  # recomputed on demand (whole grid, this GPU alone)
        return lambda: calc_alpha_line_at_nu(stellar_plasma, stellar_model, nus_q, line_cfg)[which]

    alpha_line = shard_array(L.BUF_ALPHA_LINE)
    od["alpha_line_at_nu"] = (alpha_line if alpha_line is not None
                              else DeviceArray(None, None, (D, W), fetch=lambda: _atomic(0)()[:, p0:p1]))
    if n_lines:
        stored = None if line_cfg.include_molecules else (line_table_array(L.BUF_GAMMAS), line_table_array(L.BUF_DOPPLER))
        if stored is None or stored[0] is None:
            # the molecular pass re-used the line-table buffers (or the tables live on another decomposition): atomic
            # gammas are recomputed when asked for
            od["alpha_line_at_nu_gammas"] = DeviceArray(None, None, (n_lines, D), fetch=_atomic(1))
            od["alpha_line_at_nu_doppler_widths"] = DeviceArray(None, None, (n_lines, D), fetch=_atomic(2))
        else:
            od["alpha_line_at_nu_gammas"], od["alpha_line_at_nu_doppler_widths"] = stored
    else:
        od["alpha_line_at_nu_gammas"] = np.zeros((0, D))
        od["alpha_line_at_nu_doppler_widths"] = np.zeros((0, D))
    if line_cfg.include_molecules:
        mol_alpha = shard_array(L.BUF_ALPHA_MOLECULE)
        od["molecule_alpha_line_at_nu"] = (mol_alpha if mol_alpha is not None else DeviceArray(
            None, None, (D, W), fetch=lambda: calc_molecular_alpha_line_at_nu(stellar_plasma, stellar_model, nus_q, line_cfg)[0][:, p0:p1]))
        od["molecule_alpha_line_at_nu_gammas"] = mol[0]
        od["molecule_alpha_line_at_nu_doppler_widths"] = mol[1]


def _calc_alphas_depth_sharded(stellar_plasma, stellar_model, srf, opacity_config, store_components):
    """calc_alphas of a multi-GPU run: opacity stages for this rank's depth points on the whole grid (second context),
    all-to-all of the total opacity into this rank's pixel range, where the formal solution finds it.  Per-term arrays
    are redistributed as well when ``store_components`` is set; otherwise they are recomputed on demand."""
    import torch

    from ....distributed import (DepthSlicedModel, DepthSlicedPlasma, all_shards, allgather_depth_columns, depth_indices,
                                 exchange_depth_to_nu)

    rank, world = int(srf.depth_shard[0]), int(srf.depth_shard[1])
    ctx = _ctx_of(srf)
    ctx_op = opacity_context_of(ctx)
    ctx.evict()
    ctx_op.evict()
    nus_q = srf.frequencies
    nus = u.values_of(nus_q)
    N = nus.shape[0]
    D = stellar_model.no_of_depth_points
    extent = (float(nus.min()), float(nus.max()))
    if not extent[1] <= RAYLEIGH_UPPER_BOUND_HZ:
        raise NotImplementedError("frequency grids reaching above 2.3e15 Hz are not supported (see calc_alphas)")
    bounds = getattr(srf, "shard_bounds", None) or all_shards(N, world)
    p0, p1 = _shard_of(srf, N)
    if (p0, p1) != tuple(bounds[rank]):
        raise ValueError(f"rank {rank}: pixel range {(p0, p1)} is not entry {rank} of the partition {bounds}")
    W = p1 - p0
    idx = depth_indices(D, rank, world)
    model_r = DepthSlicedModel(stellar_model, idx)
    plasma_r = DepthSlicedPlasma.of(stellar_plasma, idx)
    set_device_atmosphere(ctx_op, model_r, plasma_r)
    ctx_op.set_grid(nus)
    n_lines, mol = _device_opacity_pass(ctx_op, plasma_r, model_r, nus_q, opacity_config, store_components, collective="depth",
                                        extent=extent)
    dev = torch.device("cuda", ctx.device)
    stream = torch.cuda.current_stream(dev)

    def to_pixel_range(which):
        """(D_r, N) result of the opacity context -> (D, W) tensor on this rank's pixel range (collective)."""
        t = torch.empty((len(idx), N), dtype=torch.float64, device=dev)
        ctx_op.get(which, out=t)
        ctx_op.synchronize()            # the copy ran on the context's stream, the collective runs on torch's
        out = exchange_depth_to_nu(t, D, N, bounds=bounds)
        stream.synchronize()
        return out

    total_t = to_pixel_range(L.BUF_TOTAL)
    set_device_atmosphere(ctx, stellar_model, stellar_plasma)
    ctx.set_grid_from(ctx_op, p0, p1)  # same device: no second upload of the grid
    ctx.set_total(total_t)
    ctx.owner = getattr(srf, "token", None)

    def tensor_array(t):
        return DeviceArray(None, None, tuple(t.shape), fetch=lambda: t.cpu().numpy())

    def shard_array(which):
        if which == L.BUF_TOTAL:
            return ctx.track(DeviceArray(ctx, which, (D, W)))
        return tensor_array(to_pixel_range(which)) if store_components else None

    def line_table_array(which):
        if not store_components:
            return None
        t = torch.empty((n_lines, len(idx)), dtype=torch.float64, device=dev)
        ctx_op.get(which, out=t)
        ctx_op.synchronize()
        out = allgather_depth_columns(t, D)
        stream.synchronize()
        return tensor_array(out)

    if mol is not None and store_components:  # molecular (gammas, Doppler widths): (L, 1) gammas are depth independent
        mol = (mol[0], allgather_depth_columns(torch.from_numpy(np.ascontiguousarray(mol[1])).to(dev), D).cpu().numpy())
    _publish(srf, stellar_plasma, stellar_model, nus_q, opacity_config, store_components, n_lines, mol, (D, W), (p0, p1),
             shard_array, line_table_array)
    previous = srf.opacities._total
    total = shard_array(L.BUF_TOTAL)
    if previous is not None and np.any(np.asarray(previous) != 0):
        total = np.asarray(previous) + total.numpy()
        ctx.set_total(total)
    srf.opacities.total_alphas = total
    return total
