"""Generate tests/golden/*.npz from the reference's OWN, unmodified code (build container only).

TEST INFRASTRUCTURE.  Run as ``python -m oracle.make_golden`` from the repo root in a container that has
``/root/reference``; the reference modules are executed from where they lie through ``oracle/ref_shim.py``
(numba JIT, no source is copied).  Every fixture stores the seeded INPUTS next to the reference OUTPUTS so
that the tests never need the reference tree or this script at run time.

Fixtures
  kernels_golden.npz    faddeeva / voigt_profile / broadening scalars / weights / blackbody
  broadening_golden.npz calc_gamma + calc_doppler_width (+ VALD variants) on a synthetic line table
  alan_golden.npz       calc_alan_entries on three small problems (incl. strong lines, (L,1) gammas, NaN/inf)
  raytrace_golden.npz   single_theta_trace_parallel (plane-parallel + inward rays), raytrace() incl. spherical
  continuum_golden.npz  calc_alpha_bf/ff/rayleigh/electron/file through duck-typed plasma/model objects
  lineselect_golden.npz calc_alpha_line_at_nu (pandas merge/sort/filter + broadening + Voigt) end to end
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

from oracle.ref_shim import FakeQuantity, load_reference, REFERENCE_ROOT, _Q

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

KB = 1.380649e-16
H = 6.62607015e-27
C = 2.99792458e10
AMU = 1.66053906660e-24


def sun_atmosphere():
    """T, Pe, Pg, depth of docs/quickstart/sun.mod rows, flipped deepest -> surface (io/model/marcs.py:203-205)."""
    path = os.path.join(REFERENCE_ROOT, "docs", "quickstart", "sun.mod")
    rows = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(rows) if ln.strip().startswith("k lgTauR"))
    tab = np.array([[float(v) for v in ln.split()] for ln in rows[start + 1:start + 57]])
    depth, T, Pe, Pg = tab[:, 3], tab[:, 4], tab[:, 5], tab[:, 6]
    r = -depth[::-1]
    T = T[::-1].copy()
    n_e = (Pe / (KB * tab[:, 4]))[::-1].copy()
    n_H = (0.9 * (Pg - Pe) / (KB * tab[:, 4]))[::-1].copy()
    return r.copy(), T, n_e, n_H


def synth_lines(rng, L, nu_lo, nu_hi, T, strong_frac=0.02):
    """Seeded synthetic line table (SURVEY.md 8d recipe, small)."""
    D = T.size
    nu = np.sort(rng.uniform(nu_lo, nu_hi, L))
    Z = rng.integers(1, 31, L)
    Z[rng.random(L) < 0.05] = 1
    ion = np.where(Z == 1, 0, rng.integers(0, 2, L))
    e_ion = (rng.uniform(5.0, 25.0, L) * 1.602176634e-12)
    e_lo = rng.uniform(0.0, 0.75, L) * e_ion
    e_up = e_lo + H * nu
    bad = e_up >= e_ion
    e_ion[bad] = e_up[bad] * 1.05
    A_ul = 10.0 ** rng.uniform(6, 9, L)
    mass = (2.0 * Z + rng.uniform(0, 1, L)) * AMU
    mass[Z == 1] = 1.008 * AMU
    a0 = 10.0 ** rng.uniform(-2, 5, L)
    strong = rng.random(L) < strong_frac
    a0[strong] = 10.0 ** rng.uniform(6, 9, strong.sum())
    alpha = a0[:, None] * np.exp(-e_lo[:, None] / (KB * T[None, :])) / np.exp(-e_lo[:, None] / (KB * T.max()))
    return dict(nu=nu, atomic_number=Z.astype(np.int64), ion_number=ion.astype(np.int64), ionization_energy=e_ion,
                level_energy_lower=e_lo, level_energy_upper=e_up, A_ul=A_ul, mass=mass, alpha_line=alpha)


def gen_kernels(R, rng):
    out = {}
    # Faddeeva: all four regions + points near the region borders
    x = np.concatenate([rng.uniform(-30, 30, 3000), rng.uniform(-6, 6, 3000), np.array([0.0, 15.0, -15.0, 5.5, 5.4, 20, 7, 0.3, 0.5, 3, 2])])
    y = np.concatenate([10.0 ** rng.uniform(-6, 1.3, 3000), 10.0 ** rng.uniform(-4, 0.5, 3000), np.array([0.0, 0.0, 0.0, 0.0, 0.05, 0.01, 0.5, 0.01, 2, 0.01, 0.01])])
    z = x + 1j * y
    out["fad_z"] = z
    out["fad_w"] = R.voigt.faddeeva(z)
    dn = rng.uniform(-5e10, 5e10, 4000)
    dw = 10.0 ** rng.uniform(8.5, 10, 4000)
    g = 10.0 ** rng.uniform(6, 10.5, 4000)
    out["vp_dnu"], out["vp_dw"], out["vp_gamma"] = dn, dw, g
    out["vp_phi"] = R.voigt.voigt_profile(dn, dw, g)
    tau = np.concatenate([10.0 ** rng.uniform(-8, 3, 2000), np.array([0.0, 5e-4, 4.999e-4, 50.0, 49.999, 1e-5, 1e-3, 0.7])])
    w0, w1, w2 = R.solver.calc_weights_parallel(tau.reshape(1, -1).copy())
    out["w_tau"], out["w0"], out["w1"], out["w2"] = tau, w0[0], w1[0], w2[0]
    nus = np.linspace(1e14, 3e15, 500)
    T = np.linspace(3000, 12000, 17)
    out["bb_nus"], out["bb_T"] = nus, T
    out["bb"] = R.blackbody.blackbody_flux_at_nu(nus, T.reshape(-1, 1))
    # scalar broadening functions over random arguments
    n = 500
    nu_l = rng.uniform(2.5, 9.0, n)
    nl_l = nu_l - rng.uniform(0.2, 2.0, n)
    ne = 10.0 ** rng.uniform(10, 15, n)
    Tt = rng.uniform(3500, 12000, n)
    nH = 10.0 ** rng.uniform(13, 18, n)
    zeff = rng.integers(1, 4, n).astype(np.int64)
    out["b_nu"], out["b_nl"], out["b_ne"], out["b_T"], out["b_nH"], out["b_zeff"] = nu_l, nl_l, ne, Tt, nH, zeff
    out["b_linear_stark"] = R.broadening.calc_gamma_linear_stark(nu_l, nl_l, ne)
    out["b_quadratic_stark"] = R.broadening.calc_gamma_quadratic_stark(zeff, nu_l, nl_l, ne, Tt)
    out["b_van_der_waals"] = R.broadening.calc_gamma_van_der_waals(zeff, nu_l, nl_l, Tt, nH)
    e_ion = rng.uniform(5, 25, n) * 1.602176634e-12
    e_lev = e_ion * rng.uniform(0, 1.1, n)  # some above the ionisation energy -> NaN
    out["b_eion"], out["b_elev"] = e_ion, e_lev
    out["b_neff"] = R.broadening.calc_n_effective(zeff, e_ion, e_lev)
    nuline = rng.uniform(3e14, 1e15, n)
    mass = rng.uniform(1, 60, n) * AMU
    out["b_nuline"], out["b_mass"] = nuline, mass
    out["b_doppler"] = R.broadening.calc_doppler_width(nuline, Tt, mass, 1.3e5)
    np.savez_compressed(os.path.join(OUT, "kernels_golden.npz"), **out)


def gen_broadening(R, rng):
    r, T, n_e, n_H = sun_atmosphere()
    sel = np.arange(0, 56, 5)
    T, n_e, n_H = T[sel], n_e[sel], n_H[sel]
    lines = synth_lines(rng, 400, 4.3e14, 4.7e14, T)
    # a few auto-ionising lines (upper level above the ionisation energy): NaN n_eff
    lines["level_energy_upper"][::57] = lines["ionization_energy"][::57] * 1.01
    out = {f"line_{k}": v for k, v in lines.items()}
    out.update(T=T, n_e=n_e, n_H=n_H, vmic=1.0e5)
    args = (lines["atomic_number"][:, None], (lines["ion_number"] + 1)[:, None], lines["ionization_energy"][:, None],
            lines["level_energy_upper"][:, None], lines["level_energy_lower"][:, None], lines["A_ul"][:, None], n_e, T, n_H)
    for flags in (0, 1, 2, 4, 8, 15, 10, 5):
        out[f"gamma_{flags}"] = R.broadening.calc_gamma(*args, bool(flags & 1), bool(flags & 2), bool(flags & 4), bool(flags & 8))
    out["doppler"] = R.broadening.calc_doppler_width(lines["nu"][:, None], T, lines["mass"][:, None], 1.0e5)
    # VALD parameter variants through the reference's calc_vald_gamma (duck-typed lines/model/plasma)
    import pandas as pd

    L = len(lines["nu"])
    stark = np.where(rng.random(L) < 0.2, 0.0, -rng.uniform(4.5, 6.5, L))
    stark[::41] = 0.3  # positive: "hydrogen" marker -> 0
    kind = rng.integers(0, 4, L)
    waals = np.where(kind == 0, -rng.uniform(7.0, 8.0, L), np.where(kind == 1, 0.0, np.where(kind == 2, rng.uniform(0.5, 3.0, L), rng.integers(150, 1500, L) + rng.uniform(0.15, 0.35, L))))
    out["line_stark"], out["line_waals"] = stark, waals
    df = pd.DataFrame({k: lines[k] for k in ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_upper", "level_energy_lower", "A_ul")})
    df["stark"], df["waals"] = stark, waals
    # nuclide_masses.loc[atomic_number]: give every line its own mass through a per-line "atomic number" index
    masses = pd.Series(lines["mass"], index=np.arange(L))
    model = types.SimpleNamespace(no_of_depth_points=T.size, temperatures=FakeQuantity(T, "K"),
                                  composition=types.SimpleNamespace(nuclide_masses=types.SimpleNamespace(loc=_MassLoc(lines["mass"]))))
    plasma = types.SimpleNamespace(electron_densities=pd.Series(n_e), ion_number_density=_IonDensity({(1, 0): n_H}))
    for flags in (15, 2, 4, 9):
        out[f"vald_gamma_{flags}"] = R.broadening.calc_vald_gamma(df, model, plasma, bool(flags & 1), bool(flags & 2), bool(flags & 4), bool(flags & 8))
    np.savez_compressed(os.path.join(OUT, "broadening_golden.npz"), **out)


class _MassLoc:
    """nuclide_masses.loc[lines.atomic_number].values -> the per-line masses of the fixture."""

    def __init__(self, m):
        self.m = m

    def __getitem__(self, key):
        return types.SimpleNamespace(values=self.m)


class _IonDensity:
    def __init__(self, d):
        import pandas as pd

        self.d = {k: pd.Series(v) for k, v in d.items()}
        self.loc = self

    def __getitem__(self, key):
        return self.d[tuple(key)]


def gen_alan(R, rng):
    out = {}
    r, T, n_e, n_H = sun_atmosphere()
    cases = {
        # name: (N, lambda range, L, D selection, strong fraction)
        "a": (3000, (6550.0, 6580.0), 300, np.arange(0, 56, 8), 0.03),
        "b": (1200, (4000.0, 4012.0), 150, np.arange(0, 56, 14), 0.10),
        "c": (800, (9000.0, 9008.0), 60, np.array([0, 30, 55]), 0.0),
    }
    for name, (N, (l0, l1), L, dsel, sf) in cases.items():
        lam = np.linspace(l0, l1, N, endpoint=False)
        nus = C / (lam * 1e-8)
        Ts, nes, nHs = T[dsel], n_e[dsel], n_H[dsel]
        lines = synth_lines(rng, L, nus.min(), nus.max(), Ts, strong_frac=sf)
        gam = R.broadening.calc_gamma(lines["atomic_number"][:, None], (lines["ion_number"] + 1)[:, None],
                                      lines["ionization_energy"][:, None], lines["level_energy_upper"][:, None],
                                      lines["level_energy_lower"][:, None], lines["A_ul"][:, None], nes, Ts, nHs,
                                      True, True, True, True)
        dws = R.broadening.calc_doppler_width(lines["nu"][:, None], Ts, lines["mass"][:, None], 1.0e5)
        alph = lines["alpha_line"].copy()
        if name == "b":  # quirks: NaN gamma, infinite / >2^63 half-width, zero and negative alpha
            gam[5, :] = np.nan
            alph[9, 1] = np.inf
            alph[11, 0] = 1e40
            alph[13, :] = 0.0
            alph[17, 2] = -3.0
        if name == "c":  # radiation-only broadening: gammas of shape (L,1)
            gam = lines["A_ul"][:, None].copy()
        res = R.opac.calc_alan_entries(len(dsel), nus, lines["nu"], dws, gam, alph)
        out.update({f"{name}_nus": nus, f"{name}_line_nus": lines["nu"], f"{name}_dws": dws, f"{name}_gammas": gam,
                    f"{name}_alphas": alph, f"{name}_out": res})
    np.savez_compressed(os.path.join(OUT, "alan_golden.npz"), **out)


def gen_raytrace(R, rng):
    out = {}
    r, T, n_e, n_H = sun_atmosphere()
    N = 240
    lam = np.linspace(3000.0, 10000.0, N)
    nus = C / (lam * 1e-8)
    # continuum-like opacity rising inward + a few "lines"; some exact zeros to hit the tau == 0 branches
    base = 1e-7 * (n_H / n_H.max()) ** 0.9
    alphas = base[:, None] * (1.0 + 0.3 * np.sin(np.arange(N) / 7.0))[None, :] * (lam / 5000.0)[None, :] ** 1.5
    for c in (50, 120, 170):
        alphas[:, c - 3:c + 4] *= 10.0 ** rng.uniform(1, 4)
    alphas[:, 200] = 0.0
    alphas[40:, 201] = 0.0
    alphas[:3, 202] = 0.0
    dist = np.diff(r)
    out.update(nus=nus, T=T, alphas=alphas, r=r)
    for nth in (1, 3, 10):
        x, w = np.polynomial.legendre.leggauss(nth)
        thetas, weights = (x / 2) + 0.5 * np.pi / 2, w * np.pi / 2
        out[f"thetas_{nth}"], out[f"weights_{nth}"] = thetas, weights
    th = out["thetas_3"][1]
    out["I_single"] = R.solver.single_theta_trace_parallel(dist / np.cos(th), T.reshape(-1, 1), alphas, nus, R.blackbody.blackbody_flux_at_nu, False)
    # full raytrace(), plane-parallel and spherical, through duck-typed model / radiation field
    for tag, spherical in (("pp", False), ("sph", True)):
        rr = r - r.min() + (0.0 if not spherical else 7.0e10)
        geom = types.SimpleNamespace(r=rr, dist_to_next_depth_point=np.diff(rr), reference_r=rr[-1] - 2.0e7)
        model = types.SimpleNamespace(spherical=spherical, geometry=geom, temperatures=FakeQuantity(T, "K"))
        nth = 10 if not spherical else 4
        x, w = np.polynomial.legendre.leggauss(nth)
        srf = types.SimpleNamespace(thetas=(x / 2) + 0.5 * np.pi / 2, I_nus_weights=w * np.pi / 2,
                                    opacities=types.SimpleNamespace(total_alphas=alphas), frequencies=nus,
                                    source_function=R.blackbody.blackbody_flux_at_nu, track_individual_intensities=True,
                                    I_nus=np.zeros((T.size, N, nth)), F_nu=np.zeros((T.size, N)))
        R.solver.raytrace(model, srf)
        out[f"F_{tag}"], out[f"I_{tag}"], out[f"r_{tag}"], out[f"refr_{tag}"] = srf.F_nu, srf.I_nus, rr, geom.reference_r
        if spherical:
            out["sph_ray"] = R.solver.calculate_spherical_ray(srf.thetas, rr)
    np.savez_compressed(os.path.join(OUT, "raytrace_golden.npz"), **out)


def gen_rotation(R, rng):
    """rotation_broadening (broadening.py:824-877) on a synthetic spectrum with absorption lines, three rotational
    velocities (one below the 1e-5 km/s cut: returned unchanged)."""
    from oracle.ref_shim import UQ, Unit

    n = 4000
    lam = 5000.0 * np.exp(np.arange(n) * 2.0 / 2.99792458e5)  # constant 2 km/s per pixel
    flux = 1.0e6 * (1.0 + 0.05 * np.sin(np.arange(n) / 300.0))
    for c in rng.integers(50, n - 50, 40):
        flux *= 1.0 - rng.uniform(0.1, 0.9) * np.exp(-0.5 * ((np.arange(n) - c) / rng.uniform(1.5, 6.0)) ** 2)
    out = dict(lam=lam, flux=flux, velocity_per_pix=2.0)
    kms = Unit("km/s")
    for tag, v, eps in (("a", 17.0, 0.6), ("b", 60.5, 0.3), ("c", 0.0, 0.6)):
        _, f = R.broadening.rotation_broadening(UQ(2.0, kms), lam, flux, v_rot=UQ(v, kms), limb_darkening=eps)
        out[f"v_{tag}"], out[f"eps_{tag}"], out[f"out_{tag}"] = v, eps, np.asarray(getattr(f, "value", f), dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "rotation_golden.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = load_reference()
    if "rotation" in sys.argv[1:]:
        gen_rotation(R, np.random.default_rng(105))
        return
    gen_kernels(R, np.random.default_rng(101))
    gen_broadening(R, np.random.default_rng(102))
    gen_alan(R, np.random.default_rng(103))
    gen_raytrace(R, np.random.default_rng(104))
    gen_rotation(R, np.random.default_rng(105))
    from oracle import make_golden_pipeline

    make_golden_pipeline.main(R)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
