"""The CPU oracle (oracle/) against (1) the reference's own known-answer tests and (2) golden vectors
generated from the reference's unmodified code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden

KB = 1.380649e-16
C = 2.99792458e10


# ---- known answers of the reference's unit tests (opacities_solvers/tests/test_voigt.py, test_broadening.py)
def test_faddeeva_known_answers(oracle):
    # test_voigt.py:22-37
    assert oracle.faddeeva(0)[0] == 1 + 0j
    assert np.allclose(oracle.faddeeva(np.array([0.0, 0.0])), 1 + 0j)


def test_voigt_known_answers(oracle):
    # test_voigt.py:151-178
    assert np.allclose(oracle.voigt_profile(0, 1, 0), 1 / np.sqrt(np.pi))
    assert np.allclose(oracle.voigt_profile(0, 2, 0), 1 / (np.sqrt(np.pi) * 2))
    # test_voigt.py:130-148
    with pytest.raises(ZeroDivisionError):
        oracle.voigt_profile(0, 0, 0)


def test_broadening_known_answers(oracle):
    # test_broadening.py:40-72
    assert np.allclose(oracle.calc_doppler_width(C, 0.5, KB, 0.0), 1.0)
    # test_broadening.py:146-177
    ryd_e = 6.62607015e-27 * C * 109737.31568160
    assert np.allclose(oracle.calc_n_effective(1, ryd_e, 0.0), 1.0)
    # test_broadening.py:251-282
    assert np.allclose(oracle.calc_gamma_linear_stark(1.0, 0.0, (0.6 * 0.642) ** -1.5), 1.0)
    # test_broadening.py:355-403
    c4pref = (4.803204712570263e-10**2 * 5.29177210903e-9**3) / (36.0 * 6.62607015e-27 * (1 / (4 * np.pi)))
    ne = 1e-19 / KB * (c4pref * 36) ** (-2.0 / 3.0)
    assert np.allclose(oracle.calc_gamma_quadratic_stark(1, 1.0, 0.0, ne, 1.0), 1.0)
    # test_broadening.py:494-532
    T = np.pi * 1.67262192369e-24 / 8 / KB / 17 ** (1 / 0.3)
    assert np.allclose(oracle.calc_gamma_van_der_waals(1, 1.0, 0.0, T, (3 * 6.46e-34) ** -0.4), 1.0)


def test_survey_spot_values(oracle):
    # values produced by the reference code itself (full repr)
    w = oracle.faddeeva(np.array([20 + 0.01j, 7 + 0.5j, 0.3 + 0.01j, 0.5 + 2j, 3 + 0.01j, 2 + 0.01j, 5.4 + 0.05j]))
    exp = np.array([1.4157739178123448e-05 + 0.02824477805340991j, 0.005910408095727246 + 0.0810114188643227j,
                    0.9046221922838852 + 0.3134795946712065j, 0.24527598599538214 + 0.051521473159094625j,
                    0.0009088371132038613 + 0.20114649252072334j, 0.02062005628923175 + 0.339281612396376j,
                    0.0010219408180438072 + 0.10636210290866645j])
    np.testing.assert_allclose(w, exp, rtol=1e-13)
    np.testing.assert_allclose(oracle.voigt_profile([1.5e9, 4e10, 0], 2e9, 3e8),
                               [1.5887609343884587e-10, 1.0758574708712616e-14, 2.7372034681074e-10], rtol=1e-13)
    th, w = oracle.thetas_and_weights(3)
    np.testing.assert_allclose(th, [0.3980998287767066, 0.7853981633974483, 1.17269649801819], rtol=1e-15)
    np.testing.assert_allclose(w, [0.872664625997165, 1.3962634015954636, 0.872664625997165], rtol=1e-15)
    # tiny formal solve of SURVEY 8(c)
    T = np.array([9000, 7500, 6000, 5000, 4200.0])
    a = np.outer([1e-6, 3e-7, 8e-8, 2e-8, 5e-9], [1, 1.5, 2.5])
    I = oracle.single_theta_trace(np.array([1e6, 1.2e6, 1.5e6, 2e6]) / np.cos(0.6), T, a, np.array([6e14, 5e14, 4e14]))
    np.testing.assert_allclose(I[-1], [4.781657266276258e-05, 6.050208241034159e-05, 6.404640505720685e-05], rtol=1e-12)


# ---- golden vectors from the reference's own code
def test_kernels_golden(oracle):
    g = golden("kernels_golden.npz")
    w = oracle.faddeeva(g["fad_z"])
    np.testing.assert_allclose(w.real, g["fad_w"].real, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(w.imag, g["fad_w"].imag, rtol=1e-11, atol=1e-18)
    np.testing.assert_allclose(oracle.voigt_profile(g["vp_dnu"], g["vp_dw"], g["vp_gamma"]), g["vp_phi"], rtol=1e-12)
    w0, w1, w2 = oracle.calc_weights(g["w_tau"])
    np.testing.assert_allclose(w0, g["w0"], rtol=1e-14)
    np.testing.assert_allclose(w1, g["w1"], rtol=1e-14, atol=1e-300)
    np.testing.assert_allclose(w2, g["w2"], rtol=1e-14, atol=1e-300)
    np.testing.assert_allclose(oracle.blackbody_flux_at_nu(g["bb_nus"], g["bb_T"]), g["bb"], rtol=1e-13)
    n = g["b_nu"].size
    ls = [oracle.calc_gamma_linear_stark(g["b_nu"][i], g["b_nl"][i], g["b_ne"][i]) for i in range(n)]
    qs = [oracle.calc_gamma_quadratic_stark(g["b_zeff"][i], g["b_nu"][i], g["b_nl"][i], g["b_ne"][i], g["b_T"][i]) for i in range(n)]
    vw = [oracle.calc_gamma_van_der_waals(g["b_zeff"][i], g["b_nu"][i], g["b_nl"][i], g["b_T"][i], g["b_nH"][i]) for i in range(n)]
    ne = [oracle.calc_n_effective(g["b_zeff"][i], g["b_eion"][i], g["b_elev"][i]) for i in range(n)]
    dw = [oracle.calc_doppler_width(g["b_nuline"][i], g["b_T"][i], g["b_mass"][i], 1.3e5) for i in range(n)]
    np.testing.assert_allclose(ls, g["b_linear_stark"], rtol=1e-13)
    np.testing.assert_allclose(qs, g["b_quadratic_stark"], rtol=1e-13)
    np.testing.assert_allclose(vw, g["b_van_der_waals"], rtol=1e-13)
    np.testing.assert_allclose(ne, g["b_neff"], rtol=1e-14, equal_nan=True)
    assert np.isnan(g["b_neff"]).any()
    np.testing.assert_allclose(dw, g["b_doppler"], rtol=1e-14)


def _lines(g):
    return {k[5:]: g[k] for k in g.files if k.startswith("line_")}


def test_broadening_golden(oracle):
    g = golden("broadening_golden.npz")
    lines = _lines(g)
    for flags in (0, 1, 2, 4, 8, 15, 10, 5):
        gam, dws = oracle.calc_broadening(lines, g["T"], g["n_e"], g["n_H"], float(g["vmic"]), flags)
        np.testing.assert_allclose(gam, g[f"gamma_{flags}"], rtol=1e-13, equal_nan=True)
    np.testing.assert_allclose(dws, g["doppler"], rtol=1e-14)
    assert np.isnan(g["gamma_15"]).any()  # auto-ionising lines give NaN
    for flags in (15, 2, 4, 9):
        gam, dws = oracle.calc_broadening(lines, g["T"], g["n_e"], g["n_H"], float(g["vmic"]), flags, vald=True)
        np.testing.assert_allclose(gam, g[f"vald_gamma_{flags}"], rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_alan_entries_golden(oracle, case):
    g = golden("alan_golden.npz")
    nus, ln, dws, gam, al = (g[f"{case}_{k}"] for k in ("nus", "line_nus", "dws", "gammas", "alphas"))
    ref = g[f"{case}_out"]
    out, evals, hist = oracle.calc_alan_entries(dws.shape[1], nus, ln, dws, gam, al, with_stats=True)
    assert evals > 0 and hist.sum() == evals
    np.testing.assert_allclose(out, ref, rtol=1e-11, atol=1e-300, equal_nan=True)
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    # a nu shard equals the same columns of the full result (global idx / d_nu / window)
    p0, p1 = nus.size // 3, nus.size // 3 + 257
    shard = oracle.calc_alan_entries(dws.shape[1], nus, ln, dws, gam, al, p0=p0, p1=p1)
    np.testing.assert_allclose(shard, out[:, p0:p1], rtol=1e-13, atol=1e-300, equal_nan=True)
    # windows: pixels outside every window of a depth row stay exactly zero
    lo, hi = oracle.line_windows(nus, ln, dws, gam, al)
    cover = np.zeros(ref.shape, dtype=bool)
    for l in range(ln.size):
        for d in range(dws.shape[1]):
            cover[d, lo[l, d]:hi[l, d]] = True
    assert np.array_equal(ref != 0, cover | (ref != 0)) and not np.any(ref[~cover] != 0)


def test_raytrace_golden(oracle):
    g = golden("raytrace_golden.npz")
    nus, T, alphas, r = g["nus"], g["T"], g["alphas"], g["r"]
    for n in (1, 3, 10):
        th, w = oracle.thetas_and_weights(n)
        np.testing.assert_allclose(th, g[f"thetas_{n}"], rtol=1e-15)
        np.testing.assert_allclose(w, g[f"weights_{n}"], rtol=1e-15)
    th = g["thetas_3"][1]
    with np.errstate(divide="ignore"):
        I = oracle.single_theta_trace(np.diff(r) / np.cos(th), T, alphas, nus)
    np.testing.assert_allclose(I, g["I_single"], rtol=1e-12, atol=1e-300)
    th, w = oracle.thetas_and_weights(10)
    F, I_nus = oracle.raytrace(T, alphas, nus, th, w, dist=np.diff(g["r_pp"]), track=True)
    np.testing.assert_allclose(F, g["F_pp"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(I_nus, g["I_pp"], rtol=1e-12, atol=1e-300)
    th, w = oracle.thetas_and_weights(4)
    np.testing.assert_allclose(oracle.calculate_spherical_ray(th, g["r_sph"]), g["sph_ray"], rtol=1e-13, atol=1e-300)
    F, I_nus = oracle.raytrace(T, alphas, nus, th, w, r=g["r_sph"], spherical=True, reference_r=float(g["refr_sph"]), track=True)
    np.testing.assert_allclose(F, g["F_sph"], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(I_nus, g["I_sph"], rtol=1e-12, atol=1e-300)


# ------------------------------------------------------------------ VALD line strengths
def _vald_inputs(g):
    ll = {k[3:]: g[k] for k in g.files if k.startswith("ll_")}
    return ll


@pytest.mark.parametrize("kind", ["long", "short"])
def test_alpha_line_vald_oracle_vs_reference_golden(oracle, kind):
    """oracle.alpha_line_vald against AlphaLineVald / AlphaLineShortlistVald (plasma/base.py:178-455) run unmodified
    through oracle/ref_shim_plasma.py (tests/golden/plasma_golden.npz)."""
    g = golden("plasma_golden.npz")
    ll = _vald_inputs(g)
    alphas, lines = oracle.alpha_line_vald(ll["atomic_number"], ll["ion_charge"], ll["wavelength"], ll["log_gf"], ll["e_low"],
                                           ll["e_up"], ll["j_lo"], ll["rad"], g["ions"], g["ion_number_density"],
                                           g["partition_function"], g["T"], g["ionization_index"], g["ionization_energy"],
                                           int(g["max_atomic_number"]), shortlist=(kind == "short"))
    assert alphas.shape == g[f"{kind}_alpha"].shape
    np.testing.assert_allclose(alphas, g[f"{kind}_alpha"], rtol=1e-14)
    for col in ("atomic_number", "ion_number", "nu", "level_energy_lower", "level_energy_upper", "A_ul", "ionization_energy"):
        np.testing.assert_allclose(lines[col], g[f"{kind}_lines_{col}"], rtol=1e-14, err_msg=col)
    assert (kind == "long") == (alphas.shape[0] < (ll["atomic_number"] <= 28).sum())  # auto-ionising lines dropped (long only)


def test_molecule_plasma_restatement_vs_reference_golden(oracle):
    """oracle.molecule_number_density / molecule_partition_function / stimulated emission + AlphaLine against the
    reference's unmodified classes (plasma/molecules.py:16-191, plasma/base.py:130-175) run through ref_shim_plasma."""
    g = golden("plasma_golden.npz")
    dens, ion_map = oracle.molecule_number_density(g["mol_ion1"], g["mol_ion2"], g["mol_eq"], g["mol_t_grid"], g["mol_ion_index"],
                                                   g["mol_ion_number_density"], g["T"])
    np.testing.assert_allclose(dens, g["mol_density"], rtol=1e-13, atol=1e-300)
    np.testing.assert_array_equal(ion_map, g["mol_ion_map"])
    assert (dens[list(g["mol_names"]).index("OH-")] == 0).all() and (dens > 0).any()
    np.testing.assert_allclose(oracle.molecule_partition_function(g["mol_pf"], g["mol_t_grid"], g["T"]), g["mol_partition"], rtol=1e-15)
    np.testing.assert_allclose(oracle.alpha_line(g["al_level_number_density"], g["al_lower"], g["al_sef"], g["al_f_lu"]), g["al_alpha"], rtol=1e-14)
