"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see stardis_oracle.c header).

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may import this
module.  The product package ``stardis_b200`` never does.

Heavy loops live in ``stardis_oracle.c`` (C99 + OpenMP, loaded with ctypes); the few table-driven terms
that the reference itself delegates to numpy/scipy (``np.interp``, ``LinearNDInterpolator``) are restated
here with the same numpy/scipy calls.  Reference citations are relative to ``/root/reference/stardis/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstardis_oracle.so")

# CODATA-2018 CGS (astropy 6.1), identical to the #defines in stardis_oracle.c
C_CGS = 2.99792458e10
H_CGS = 6.62607015e-27
KB_CGS = 1.380649e-16
E_ESU = 4.803204712570263e-10
ME_CGS = 9.1093837015e-28
RYD_CGS = 109737.31568160
SIGMA_T = 6.6524587321e-25
# opacities_solvers/base.py:21-34
BF_CONSTANT = 64 * np.pi**4 * E_ESU**10 * ME_CGS / (3 * np.sqrt(3) * C_CGS * H_CGS**6)
FF_CONSTANT = 4 / (3 * H_CGS * C_CGS) * E_ESU**6 * np.sqrt(2 * np.pi / (3 * ME_CGS**3 * KB_CGS))
RYDBERG_FREQUENCY = C_CGS * RYD_CGS

LINEAR_STARK, QUADRATIC_STARK, VAN_DER_WAALS, RADIATION = 1, 2, 4, 8


def build(force: bool = False) -> str:
    """Compile libstardis_oracle.so with the committed Makefile (no-op when up to date)."""
    src = os.path.join(_HERE, "stardis_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libstardis_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.sdo_max_threads.restype = C.c_int
        L.sdo_d_nu.restype = C.c_double
        L.sdo_calc_alan_entries.restype = C.c_int64
        for n in ("sdo_doppler_width", "sdo_n_effective", "sdo_gamma_linear_stark", "sdo_gamma_quadratic_stark",
                  "sdo_gamma_van_der_waals"):
            getattr(L, n).restype = C.c_double
        L.sdo_doppler_width.argtypes = [C.c_double] * 4
        L.sdo_n_effective.argtypes = [C.c_double] * 3
        L.sdo_gamma_linear_stark.argtypes = [C.c_double] * 3
        L.sdo_gamma_quadratic_stark.argtypes = [C.c_double] * 5
        L.sdo_gamma_van_der_waals.argtypes = [C.c_double] * 5
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_ip)


def max_threads() -> int:
    return int(lib().sdo_max_threads())


def set_threads(n: int) -> None:
    lib().sdo_set_threads(int(n))


# ------------------------------------------------------------------ Voigt ----
def faddeeva(z):
    """voigt.py:17-91."""
    z = np.atleast_1d(np.asarray(z, dtype=np.complex128))
    zr, pzr = _d(z.real)
    zi, pzi = _d(z.imag)
    wr = np.empty_like(zr)
    wi = np.empty_like(zr)
    lib().sdo_faddeeva(C.c_int64(zr.size), pzr, pzi, wr.ctypes.data_as(_dp), wi.ctypes.data_as(_dp))
    return (wr + 1j * wi).reshape(z.shape)


def voigt_profile(delta_nu, doppler_width, gamma):
    """voigt.py:113-155; a zero Doppler width raises ZeroDivisionError (test_voigt.py:130-148)."""
    dn, dw, g = np.broadcast_arrays(
        np.asarray(delta_nu, dtype=np.float64), np.asarray(doppler_width, dtype=np.float64), np.asarray(gamma, dtype=np.float64)
    )
    if np.any(dw == 0):
        raise ZeroDivisionError("division by zero")
    shape = dn.shape
    dn, pdn = _d(dn.ravel())
    dw, pdw = _d(dw.ravel())
    g, pg = _d(g.ravel())
    out = np.empty_like(dn)
    lib().sdo_voigt_profile(C.c_int64(dn.size), pdn, pdw, pg, out.ctypes.data_as(_dp))
    return out.reshape(shape)


# ------------------------------------------------------------- broadening ----
def calc_doppler_width(nu_line, temperature, atomic_mass, microturbulence=0.0):
    return lib().sdo_doppler_width(nu_line, temperature, atomic_mass, microturbulence)


def calc_n_effective(ion_number, ionization_energy, level_energy):
    return lib().sdo_n_effective(float(ion_number), ionization_energy, level_energy)


def calc_gamma_linear_stark(n_eff_upper, n_eff_lower, electron_density):
    return lib().sdo_gamma_linear_stark(n_eff_upper, n_eff_lower, electron_density)


def calc_gamma_quadratic_stark(ion_number, n_eff_upper, n_eff_lower, electron_density, temperature):
    return lib().sdo_gamma_quadratic_stark(float(ion_number), n_eff_upper, n_eff_lower, electron_density, temperature)


def calc_gamma_van_der_waals(ion_number, n_eff_upper, n_eff_lower, temperature, h_density):
    return lib().sdo_gamma_van_der_waals(float(ion_number), n_eff_upper, n_eff_lower, temperature, h_density)


def calc_broadening(lines, T, n_e, n_H, vmic, flags, vald=False):
    """(gammas, doppler_widths), both (L,D).  ``lines`` is a dict of per-line columns:
    nu, atomic_number, ion_number, ionization_energy, level_energy_upper, level_energy_lower, A_ul, mass
    (+ stark, waals when ``vald``).  broadening.py:659-732 and :1009-1085."""
    L = len(lines["nu"])
    T, pT = _d(T)
    D = T.size
    ne, pne = _d(n_e)
    nH, pnH = _d(n_H)
    nu, pnu = _d(lines["nu"])
    Z, pZ = _i(lines["atomic_number"])
    ion, pion = _i(lines["ion_number"])
    ei, pei = _d(lines["ionization_energy"])
    eu, peu = _d(lines["level_energy_upper"])
    el, pel = _d(lines["level_energy_lower"])
    A, pA = _d(lines["A_ul"])
    m, pm = _d(lines["mass"])
    gam = np.empty((L, D))
    dws = np.empty((L, D))
    if vald:
        st, pst = _d(lines["stark"])
        wa, pwa = _d(lines["waals"])
        lib().sdo_calc_vald_broadening(
            C.c_int64(L), C.c_int(D), pnu, pZ, pion, pei, peu, pel, pA, pm, pst, pwa, pT, pne, pnH,
            C.c_double(vmic), C.c_uint(flags), gam.ctypes.data_as(_dp), dws.ctypes.data_as(_dp))
    else:
        lib().sdo_calc_broadening(
            C.c_int64(L), C.c_int(D), pnu, pZ, pion, pei, peu, pel, pA, pm, pT, pne, pnH,
            C.c_double(vmic), C.c_uint(flags), gam.ctypes.data_as(_dp), dws.ctypes.data_as(_dp))
    return gam, dws


# ---------------------------------------------------------- line opacity ----
def d_nu(nus):
    nus, p = _d(nus)
    return lib().sdo_d_nu(C.c_int64(nus.size), p)


def line_windows(nus, line_nus, doppler_widths, gammas, alphas):
    """[lo, hi) per (line, depth): opacities_solvers/base.py:556-575."""
    nus, pn = _d(nus)
    ln, pl = _d(line_nus)
    dw, pdw = _d(doppler_widths)
    g, pg = _d(gammas)
    a, pa = _d(alphas)
    L, D = dw.shape
    lo = np.empty((L, D), dtype=np.int64)
    hi = np.empty((L, D), dtype=np.int64)
    lib().sdo_line_windows(C.c_int64(L), C.c_int(D), C.c_int64(nus.size), pn, pl, pdw, pg, C.c_int(g.shape[1]), pa,
                           lo.ctypes.data_as(_ip), hi.ctypes.data_as(_ip))
    return lo, hi


def calc_alan_entries(no_of_depth_points, tracing_nus_values, line_nus, doppler_widths, gammas, alphas_array,
                      p0=0, p1=None, with_stats=False):
    """opacities_solvers/base.py:487-592.  Returns alpha_line_at_nu (D, p1-p0) [and (evals, region_hist)]."""
    nus, pn = _d(tracing_nus_values)
    N = nus.size
    p1 = N if p1 is None else int(p1)
    ln, pl = _d(line_nus)
    dw, pdw = _d(doppler_widths)
    g, pg = _d(gammas)
    if g.ndim == 1:
        g = g[:, None]
    a, pa = _d(alphas_array)
    D = int(no_of_depth_points)
    out = np.zeros((D, p1 - p0))
    hist = np.zeros(4, dtype=np.int64)
    ev = lib().sdo_calc_alan_entries(
        C.c_int(D), C.c_int64(N), pn, C.c_int64(ln.size), pl, pdw, pg, C.c_int(g.shape[1]), pa,
        C.c_int64(p0), C.c_int64(p1), out.ctypes.data_as(_dp), hist.ctypes.data_as(_ip) if with_stats else None)
    if ev < 0:
        raise MemoryError("oracle slab allocation failed")
    return (out, int(ev), hist) if with_stats else out


# -------------------------------------------------------------- continuum ----
def alpha_bf(nus, nu_cut, zeff, n_level):
    """calc_alpha_bf / calc_contribution_bf (opacities_solvers/base.py:178-271).  n_level is (n_levels, D)."""
    nus, pn = _d(nus)
    nc, pc = _d(nu_cut)
    z, pz = _d(zeff)
    nl, pnl = _d(n_level)
    D = nl.shape[1] if nl.ndim == 2 else 0
    out = np.zeros((D, nus.size))
    lib().sdo_alpha_bf(C.c_int(D), C.c_int64(nus.size), pn, C.c_int(nc.size), pc, pz, pnl, C.c_double(BF_CONSTANT),
                       C.c_double(RYDBERG_FREQUENCY), out.ctypes.data_as(_dp))
    return out


def alpha_ff(nus, species, T):
    """calc_alpha_ff (opacities_solvers/base.py:274-317).  species = list of (Z_charge, n_e*n_ion[D])."""
    nus, pn = _d(nus)
    T = np.asarray(T, dtype=np.float64)
    coef = np.zeros(T.size)
    for zc, dens in species:
        coef += (np.asarray(dens) / np.sqrt(T)) * (FF_CONSTANT * zc**2)
    coef, pc = _d(coef)
    out = np.zeros((T.size, nus.size))
    lib().sdo_alpha_ff(C.c_int(T.size), C.c_int64(nus.size), pn, pc, out.ctypes.data_as(_dp))
    return out


def rayleigh_clip(nus):
    """opacities_solvers/base.py:98-99: frequencies above 2.3e15 Hz are zeroed (in place in the reference)."""
    nus = np.array(nus, dtype=np.float64)
    nus[nus > 2.3e15] = 0
    return nus


def alpha_rayleigh(nus, n_HI, n_HeI, n_H2, species):
    """calc_alpha_rayleigh (opacities_solvers/base.py:74-135)."""
    nus, pn = _d(rayleigh_clip(nus))
    D = len(n_HI)
    c4 = np.zeros(D); c6 = np.zeros(D); c8 = np.zeros(D)
    if "H" in species:
        c4 += 20.24 * n_HI; c6 += 239.2 * n_HI; c8 += 2256 * n_HI
    if "He" in species:
        c4 += 1.913 * n_HeI; c6 += 4.52 * n_HeI; c8 += 7.90 * n_HeI
    if "H2" in species:
        c4 += 28.39 * n_H2; c6 += 215.0 * n_H2; c8 += 1303 * n_H2
    out = np.zeros((D, nus.size))
    lib().sdo_alpha_rayleigh(C.c_int(D), C.c_int64(nus.size), pn, _d(c4)[1], _d(c6)[1], _d(c8)[1],
                             C.c_double(RYDBERG_FREQUENCY), C.c_double(SIGMA_T), out.ctypes.data_as(_dp))
    return out


def alpha_electron(n_e, N):
    """calc_alpha_electron (opacities_solvers/base.py:139-174)."""
    return np.repeat((SIGMA_T * np.asarray(n_e, dtype=np.float64))[:, None], N, axis=1)


def _read_table_2d(fpath, source):
    """Table ingestion of util.py:35-47 (H2plus_bf) and :66-72 (Hminus_ff), without pandas."""
    rows = [ln for ln in open(fpath).read().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    if source == "H2plus_bf":
        header = rows[0].split()[1:]
        ys = np.array([int(float(h)) for h in header], dtype=np.float64)  # temperatures
        xs, vals = [], []
        for ln in rows[1:]:
            tok = ln.split()
            xs.append(float(tok[0]))
            vals.append([float(_stancil(t)) for t in tok[1:]])
        return np.array(xs) * 10.0, ys, np.array(vals)  # nm -> Angstrom
    if source == "Hminus_ff":
        header = [h.strip(",") for h in rows[0].split()]
        header = [h for h in header if h]
        ys = np.array([float(h) for h in header])  # theta = 5040/T
        xs, vals = [], []
        for ln in rows[1:]:
            tok = ln.split()
            xs.append(float(tok[0]))
            vals.append([float(t) for t in tok[1:]])
        return np.array(xs), ys, np.array(vals)
    raise ValueError(source)


def _stancil(tok):
    """'7.34-5' -> 7.34e-5 (util.py:41 replaces '-' by 'e-')."""
    return tok.replace("-", "e-") if "-" in tok[1:] and "e" not in tok.lower() else tok


def sigma_file(tracing_lambdas, temperatures, fpath, opacity_source=None):
    """util.py:14-108, restated with the same numpy/scipy calls the reference makes."""
    from scipy.interpolate import LinearNDInterpolator

    tracing_lambdas = np.asarray(tracing_lambdas, dtype=np.float64)
    temperatures = np.asarray(temperatures, dtype=np.float64)
    if opacity_source == "H2plus_bf":
        xs, ys, vals = _read_table_2d(fpath, opacity_source)
        xm, ym = np.meshgrid(xs, ys, indexing="ij")
        interp = LinearNDInterpolator(np.vstack([xm.ravel(), ym.ravel()]).T, vals.flatten(), fill_value=0)
        lam, tt = np.meshgrid(tracing_lambdas, temperatures)
        return interp(lam, tt) * 1e-18
    if opacity_source == "Hminus_ff":
        xs, ys, vals = _read_table_2d(fpath, opacity_source)
        xm, ym = np.meshgrid(xs, ys, indexing="ij")
        interp = LinearNDInterpolator(np.vstack([xm.ravel(), ym.ravel()]).T, vals.flatten(), fill_value=0)
        lam, th = np.meshgrid(tracing_lambdas, 5040 / temperatures)
        return interp(lam, th) * 1e-26 * KB_CGS * temperatures[:, None]
    if opacity_source == "Hminus_bf":
        tab = np.array([[float(v) for v in ln.split(",")] for ln in open(fpath).read().splitlines()
                        if ln.strip() and not ln.lstrip().startswith("#")])
        return np.interp(tracing_lambdas, tab[:, 0], tab[:, 1])
    raise ValueError(f"Unknown opacity_source: {opacity_source}")


def alpha_file(nus, temperatures, fpath, source, number_density):
    """calc_alpha_file (opacities_solvers/base.py:40-70)."""
    lambdas = C_CGS / np.asarray(nus, dtype=np.float64) * 1e8
    sig = sigma_file(lambdas, temperatures, fpath, source)
    return sig * np.asarray(number_density, dtype=np.float64)[:, None]


# ----------------------------------------------------------- formal solver ----
def blackbody_flux_at_nu(nus, T):
    """source_functions/blackbody.py:11-35 -> (D,N)."""
    nus, pn = _d(nus)
    T, pT = _d(np.ravel(T))
    out = np.empty((T.size, nus.size))
    lib().sdo_blackbody(C.c_int(T.size), C.c_int64(nus.size), pn, pT, out.ctypes.data_as(_dp))
    return out


def calc_weights(tau):
    """radiation_field_solvers/base.py:6-47."""
    tau = np.asarray(tau, dtype=np.float64)
    t, pt = _d(tau.ravel())
    w = [np.empty_like(t) for _ in range(3)]
    lib().sdo_calc_weights(C.c_int64(t.size), pt, *[x.ctypes.data_as(_dp) for x in w])
    return tuple(x.reshape(tau.shape) for x in w)


def thetas_and_weights(n):
    """radiation_field/base.py:60-63 (NB: not the usual affine map onto [0, pi/2])."""
    x, w = np.polynomial.legendre.leggauss(int(n))
    return (x / 2) + 0.5 * np.pi / 2, w * np.pi / 2


def single_theta_trace(ray_ds, T, alphas, nus, inward_rays=False):
    """radiation_field_solvers/base.py:85-268 -> I (D,N)."""
    ds, pds = _d(ray_ds)
    T, pT = _d(np.ravel(T))
    a, pa = _d(alphas)
    nus, pn = _d(nus)
    D, N = a.shape
    out = np.zeros((D, N))
    lib().sdo_single_theta_trace(C.c_int(D), C.c_int64(N), pds, pT, pa, pn, C.c_int(int(inward_rays)), out.ctypes.data_as(_dp))
    return out


def calculate_spherical_ray(thetas, r):
    """radiation_field_solvers/base.py:349-381 -> (G, n_theta)."""
    th, pth = _d(thetas)
    r, pr = _d(r)
    out = np.zeros((r.size - 1, th.size))
    lib().sdo_spherical_ray(C.c_int(r.size), C.c_int(th.size), pth, pr, out.ctypes.data_as(_dp))
    return out


def raytrace(T, alphas, nus, thetas, weights, dist=None, r=None, spherical=False, reference_r=None,
             F_nu=None, track=False):
    """raytrace (radiation_field_solvers/base.py:271-346).  Returns (F_nu (D,N), I_nus (D,N,n_theta) or None)."""
    T = np.asarray(T, dtype=np.float64).ravel()
    thetas = np.asarray(thetas, dtype=np.float64)
    if spherical:
        ray_ds = calculate_spherical_ray(thetas, r)
        scale = (r[-1] / reference_r) ** 2
    else:
        ray_ds = np.asarray(dist, dtype=np.float64).reshape(-1, 1) / np.cos(thetas)
        scale = 1.0
    a, pa = _d(alphas)
    D, N = a.shape
    nus, pn = _d(nus)
    w, pw = _d(weights)
    ds, pds = _d(ray_ds)
    F = np.zeros((D, N)) if F_nu is None else np.ascontiguousarray(F_nu, dtype=np.float64)
    I_nus = np.zeros((D, N, thetas.size)) if track else None
    lib().sdo_raytrace(C.c_int(D), C.c_int64(N), C.c_int(thetas.size), pds, pw, _d(T)[1], pa, pn,
                       C.c_int(int(spherical)), C.c_double(scale), F.ctypes.data_as(_dp),
                       I_nus.ctypes.data_as(_dp) if track else None)
    return F, I_nus


# ------------------------------------------------------------------ VALD line strengths (upstream of K1/K2)
def alpha_line_vald(atomic_number, ion_charge, wavelength_aa, log_gf, e_low_ev, e_up_ev, j_lo, rad, ions, ion_number_density,
                    partition_function, t_electrons, ionization_index, ionization_energy, max_atomic_number, shortlist=False):
    """numpy restatement of ``AlphaLineVald.calculate`` (stardis/plasma/base.py:203-321) and, with ``shortlist=True``,
    of ``AlphaLineShortlistVald.calculate`` (:346-455), in the reference's order of operations.

    ``ions``: (n_ions, 2) (atomic_number, charge) rows of ``ion_number_density`` / ``partition_function`` (n_ions, D);
    ``ionization_index``: (n, 2) (atomic_number, ion_number = charge + 1) rows of ``ionization_energy`` [erg].
    Returns (alphas (L', D), lines dict) after the truncation to ``max_atomic_number`` (:235-237 / :375-377) and, for the
    long list, the removal of auto-ionising lines (:316-318)."""
    EV, KB, H, C_ = 1.602176634e-12, 1.380649e-16, 6.62607015e-27, 2.99792458e10
    ALPHA_COEFFICIENT = (np.pi * 4.803204712570263e-10 ** 2) / (9.1093837015e-28 * C_)  # :35
    keep = np.asarray(atomic_number) <= max_atomic_number
    Z, q = np.asarray(atomic_number)[keep], np.asarray(ion_charge)[keep]
    lam, lgf, e_low = np.asarray(wavelength_aa, float)[keep], np.asarray(log_gf, float)[keep], np.asarray(e_low_ev, float)[keep]
    T = np.asarray(t_electrons, float)
    if shortlist:  # :380-387
        e_up = (e_low * EV + (H * C_) / (lam * 1e-8)) / EV
    else:
        e_up = np.asarray(e_up_ev, float)[keep]
    exponent = np.exp(np.outer(-e_low * EV, 1.0 / (T * KB)))                       # :242-247 / :389-393
    row = {(int(a), int(b)): i for i, (a, b) in enumerate(np.asarray(ions))}
    n_over_u = (np.asarray(ion_number_density, float) / np.asarray(partition_function, float))[[row[(int(a), int(b))] for a, b in zip(Z, q)]]
    line_nus = C_ / (lam * 1e-8)                                                    # :268-270
    emission = 1.0 - np.exp((-H / KB) * np.outer(line_nus, 1.0 / T))               # :272-281
    if shortlist:
        prefactor = (exponent * n_over_u).T                                         # :401-403
        alphas = (ALPHA_COEFFICIENT * prefactor * 10 ** lgf * emission.T).T         # :420-427
    else:
        g_lo = np.asarray(j_lo, float)[keep] * 2 + 1                                # :240
        n_lower = (exponent * n_over_u).T * g_lo                                    # :256-262
        f_lu = 10 ** lgf / g_lo                                                     # :264-266
        alphas = (ALPHA_COEFFICIENT * n_lower * f_lu * emission.T).T                # :283-291
    ion_e = {(int(a), int(b) - 1): e for (a, b), e in zip(np.asarray(ionization_index), np.asarray(ionization_energy, float))}
    lines = dict(atomic_number=Z, ion_number=q, nu=line_nus, level_energy_lower=e_low * EV, level_energy_upper=e_up * EV,
                 A_ul=10 ** np.asarray(rad, float)[keep],
                 ionization_energy=np.array([ion_e.get((int(a), int(b)), np.nan) for a, b in zip(Z, q)]))
    if not shortlist:
        valid = lines["level_energy_upper"] < lines["ionization_energy"]
        alphas = alphas[valid]
        lines = {k: v[valid] for k, v in lines.items()}
    return alphas, lines


# ------------------------------------------------------------------ tardis line strengths (upstream of K1/K2)
def stimulated_emission_factor(level_number_density, g, lower_level_index, upper_level_index, metastable_upper):
    """tardis ``StimulatedEmissionFactor.calculate`` (third-party tardis release-2024.08.25,
    tardis/plasma/properties/radiative_properties.py; source absent offline -- restated from its published algorithm, PARITY
    UNPINNED for this factor): 1 - (g_lower n_upper) / (g_upper n_lower); 0 where n_lower == 0, where the factor is -inf, and
    where it is negative for a line whose upper level is metastable."""
    n = np.asarray(level_number_density, dtype=np.float64)
    g = np.asarray(g, dtype=np.float64)
    n_lower, n_upper = n[lower_level_index], n[upper_level_index]
    with np.errstate(divide="ignore", invalid="ignore"):
        sef = 1.0 - ((g[lower_level_index][:, None] * n_upper) / (g[upper_level_index][:, None] * n_lower))
    sef[n_lower == 0.0] = 0.0
    sef[np.isneginf(sef)] = 0.0
    sef[np.asarray(metastable_upper, dtype=bool)[:, None] & (sef < 0)] = 0.0
    return sef


def alpha_line(level_number_density, lower_level_index, stimulated_emission_factor_, f_lu):
    """``AlphaLine.calculate`` (stardis/plasma/base.py:143-158): ALPHA_COEFFICIENT * n_lower * sef * f_lu, left to right."""
    ALPHA_COEFFICIENT = (np.pi * 4.803204712570263e-10 ** 2) / (9.1093837015e-28 * 2.99792458e10)  # :35
    n_lower = np.asarray(level_number_density, dtype=np.float64)[lower_level_index]
    return ALPHA_COEFFICIENT * n_lower * stimulated_emission_factor_ * np.asarray(f_lu, dtype=np.float64)[:, None]


# ------------------------------------------------------------------ molecules (stardis/plasma/molecules.py)
_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn").split()


def split_ion(s):
    """'C+' -> (6, 1), 'H-' -> (1, -1) (molecules.py:145-159: symbol, count of '+' minus count of '-')."""
    import re

    m = re.match(r"([A-Z][a-z]?)(\+*)(\-*)", s)
    return _SYMBOLS.index(m.group(1)) + 1, len(m.group(2)) - len(m.group(3))


def molecule_number_density(ion1, ion2, equilibrium_constants, t_grid, ion_index, ion_number_density, t_electrons):
    """``MoleculeIonNumberDensity.calculate`` (molecules.py:35-143): Barklem & Collet 2016 pressure equilibrium constants
    (log10, SI) splined in T, converted to number-density constants with the ideal gas law, then the closed-form solution of
    the dissociation equilibrium (homonuclear / heteronuclear); negative ions and absent elements give 0."""
    from scipy.interpolate import CubicSpline

    T = np.asarray(t_electrons, dtype=np.float64)
    nd = {tuple(int(v) for v in k): np.asarray(r, dtype=np.float64) for k, r in zip(ion_index, ion_number_density)}
    included = {k[0] for k in nd}
    out = np.zeros((len(ion1), T.size))
    ion_map = np.zeros((len(ion1), 2), dtype=np.int64)
    for i, (a, b) in enumerate(zip(ion1, ion2)):
        (z1, c1), (z2, c2) = split_ion(str(a)), split_ion(str(b))
        ion_map[i] = (z1, z2)
        if c1 == -1 or c2 == -1 or z1 not in included or z2 not in included:
            continue
        n1, n2 = nd[(z1, c1)], nd[(z2, c2)]
        logk = CubicSpline(t_grid, equilibrium_constants[i], extrapolate=True)(T)
        k = (10.0 ** logk) * 10.0 / (1.380649e-16 * T)  # Pa -> dyn cm^-2, / k_B T -> cm^-3
        if z1 == z2 and c1 == c2:
            n = (1 / 8) * ((-((k * (k + 8 * n1)) ** 0.5)) + k + 4 * n1)
        else:
            n = 0.5 * (-np.sqrt(k ** 2 + 2 * k * (n1 + n2) + (n1 - n2) ** 2) + k + n1 + n2)
        out[i] = np.maximum(n, 0)
    return out, ion_map


def molecule_partition_function(partition_functions, t_grid, t_electrons):
    """``MoleculePartitionFunction.calculate`` (molecules.py:176-191): np.interp per molecule."""
    return np.array([np.interp(t_electrons, t_grid, row) for row in partition_functions])
