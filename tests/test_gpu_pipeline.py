"""GPU parity tests, API level: the drop-in Python API (calc_alphas / raytrace / RadiationField / run_stardis) against
pipeline goldens produced by the reference's own, unmodified calc_alphas + raytrace on the same seeded inputs."""
import os

import numpy as np
import pytest

from conftest import golden
from helpers import write_marcs_mod, write_table_files

pytestmark = pytest.mark.gpu

RTOL_ALPHA = 1e-8  # north star: alpha[depth, nu]
RTOL_F = 1e-6      # north star: emergent F_nu


@pytest.fixture(scope="module")
def table_paths(tmp_path_factory):
    return write_table_files(str(tmp_path_factory.mktemp("tables")))


def _cases():
    from oracle.make_golden_pipeline import CASES

    return list(CASES)


def _run_case(name, table_paths, hide_line_table=False):
    from oracle.make_golden_pipeline import CASES, case_inputs
    from stardis_b200 import units as u
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.opacities.opacities_solvers import calc_alphas
    from stardis_b200.radiation_field.radiation_field_solvers import raytrace
    from stardis_b200.radiation_field.source_functions.blackbody import blackbody_flux_at_nu

    cfg, model, plasma, nus = case_inputs(name, CASES[name], table_paths)
    if hide_line_table:  # force the pandas adapter (ColumnarLines.from_plasma), as with a real tardis plasma
        plasma.expose_columnar = False
    srf = RadiationField(u.Quantity(nus, u.Hz), blackbody_flux_at_nu, model, cfg.no_of_thetas, track_individual_intensities=True)
    total = calc_alphas(plasma, model, srf, cfg.opacity)
    F = raytrace(model, srf)
    return cfg, model, plasma, nus, srf, total, F


def _check_case(name, g, plasma, nus, srf, total, F):
    from oracle.make_golden_pipeline import DEPTH_ROWS

    fp = g[f"{name}__fingerprint"]
    lt = plasma._line_table
    np.testing.assert_allclose([lt.nu.sum(), lt.alpha_line.sum(), plasma.electron_densities.values.sum(), nus.sum()], fp, rtol=1e-14)
    keys = [k[len(name) + 2:] for k in g.files if k.startswith(name + "__")]
    od = srf.opacities.opacities_dict
    ref_keys = [k for k in keys if k not in ("total", "F_nu", "I_nus_emergent", "fingerprint")]
    assert list(od.keys()) == ref_keys  # same keys, same order as the reference
    for k in ref_keys:
        ref = g[f"{name}__{k}"]
        got = np.asarray(od[k], dtype=np.float64)
        if ref.ndim == 2 and ref.shape[1] == len(nus):
            got = got[DEPTH_ROWS]
        assert got.shape == ref.shape, k
        np.testing.assert_allclose(got, ref, rtol=RTOL_ALPHA, atol=1e-300, equal_nan=True, err_msg=k)
    np.testing.assert_allclose(np.asarray(total), g[f"{name}__total"], rtol=RTOL_ALPHA, equal_nan=True)
    np.testing.assert_allclose(np.asarray(srf.opacities.total_alphas), g[f"{name}__total"], rtol=RTOL_ALPHA, equal_nan=True)
    np.testing.assert_allclose(np.asarray(F), g[f"{name}__F_nu"], rtol=RTOL_F, equal_nan=True)
    np.testing.assert_allclose(np.asarray(srf.I_nus)[-1], g[f"{name}__I_nus_emergent"], rtol=RTOL_F, equal_nan=True)


@pytest.mark.parametrize("name", ["bench", "broadening", "plain", "vald", "nolines"])
def test_calc_alphas_raytrace_vs_reference(name, table_paths):
    g = golden("pipeline_golden.npz")
    cfg, model, plasma, nus, srf, total, F = _run_case(name, table_paths)
    _check_case(name, g, plasma, nus, srf, total, F)


def test_pandas_adapter_matches_reference(table_paths):
    """A plasma that only exposes the tardis-style pandas tables goes through ColumnarLines.from_plasma (the
    reference's merges/sort/filter done once) and must give the same result."""
    g = golden("pipeline_golden.npz")
    for name in ("bench", "vald"):
        cfg, model, plasma, nus, srf, total, F = _run_case(name, table_paths, hide_line_table=True)
        _check_case(name, g, plasma, nus, srf, total, F)


def test_standalone_functions_match_driver(table_paths):
    """calc_alpha_bf/ff/rayleigh/electron/file/line_at_nu called on their own (reference signatures) give the
    entries the driver produced; the lazily recomputed entries (store_components=False) as well."""
    from oracle.make_golden_pipeline import CASES, case_inputs
    from stardis_b200 import units as u
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.opacities.opacities_solvers import base as ob

    cfg, model, plasma, nus = case_inputs("broadening", CASES["broadening"], table_paths)
    q = u.Quantity(nus, u.Hz)
    srf = RadiationField(q, None, model, 3)
    ob.calc_alphas(plasma, model, srf, cfg.opacity)
    od = {k: np.array(np.asarray(v)) for k, v in srf.opacities.opacities_dict.items()}
    np.testing.assert_array_equal(ob.calc_alpha_bf(plasma, model, q, cfg.opacity.bf), od["alpha_bf"])
    np.testing.assert_array_equal(ob.calc_alpha_ff(plasma, model, q, cfg.opacity.ff), od["alpha_ff"])
    np.testing.assert_array_equal(ob.calc_alpha_rayleigh(plasma, model, q, cfg.opacity.rayleigh), od["alpha_rayleigh"])
    np.testing.assert_array_equal(ob.calc_alpha_electron(plasma, model, q), od["alpha_electron"])
    for src, path in cfg.opacity.file.items():
        np.testing.assert_array_equal(ob.calc_alpha_file(plasma, model, q, src, path), od[f"alpha_file_{src}"])
    a, gam, dws = ob.calc_alpha_line_at_nu(plasma, model, q, cfg.opacity.line)
    np.testing.assert_array_equal(a, od["alpha_line_at_nu"])
    np.testing.assert_array_equal(gam, od["alpha_line_at_nu_gammas"])
    assert gam.shape == dws.shape and len(a) == model.no_of_depth_points  # test_opacities_solvers.py:4-17
    assert ob.calc_alpha_electron(plasma, model, q, True) == 0
    with pytest.raises(ValueError):  # util.py:105-106
        ob.calc_alpha_file(plasma, model, q, "Heminus_ff", table_paths["Hminus_ff"])
    srf2 = RadiationField(q, None, model, 3)
    ob.calc_alphas(plasma, model, srf2, cfg.opacity, store_components=False)
    for k in ("alpha_bf", "alpha_file_Hminus_ff", "alpha_rayleigh"):
        np.testing.assert_array_equal(np.asarray(srf2.opacities.opacities_dict[k]), od[k])
    np.testing.assert_array_equal(np.asarray(srf2.opacities.total_alphas), np.asarray(srf.opacities.total_alphas))


def test_calc_alan_entries_api_and_zero_doppler():
    from stardis_b200.radiation_field.opacities.opacities_solvers import calc_alan_entries

    g = golden("alan_golden.npz")
    out = calc_alan_entries(g["a_dws"].shape[1], g["a_nus"], g["a_line_nus"], g["a_dws"], g["a_gammas"], g["a_alphas"])
    np.testing.assert_allclose(out, g["a_out"], rtol=RTOL_ALPHA, atol=1e-300)
    dws = g["a_dws"].copy()
    dws[3, 2] = 0.0
    with pytest.raises(ZeroDivisionError):
        calc_alan_entries(dws.shape[1], g["a_nus"], g["a_line_nus"], dws, g["a_gammas"], g["a_alphas"])


def test_raytrace_on_hand_filled_field_and_context_reuse(table_paths):
    """raytrace() on a RadiationField whose total_alphas were set by hand (as the reference's fixtures do), and
    results that survive the context being reused by a later computation."""
    from stardis_b200 import units as u
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.radiation_field_solvers import raytrace
    from stardis_b200.synthetic import load_atmosphere, stellar_model_from_atmosphere

    g = golden("raytrace_golden.npz")
    model = stellar_model_from_atmosphere(load_atmosphere("sun"))
    model.geometry.r = u.Quantity(g["r_pp"], u.cm)
    srf = RadiationField(u.Quantity(g["nus"], u.Hz), None, model, 10, track_individual_intensities=True)
    srf.opacities.total_alphas = g["alphas"].copy()
    F = raytrace(model, srf)
    srf_b = RadiationField(u.Quantity(g["nus"][:50], u.Hz), None, model, 2)
    srf_b.opacities.total_alphas = np.ascontiguousarray(g["alphas"][:, :50])
    raytrace(model, srf_b)  # reuses the context: F must have been moved to the host intact
    np.testing.assert_allclose(np.asarray(F), g["F_pp"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(np.asarray(srf.I_nus), g["I_pp"], rtol=1e-9, atol=1e-300)
    # F_nu accumulates (radiation_field_solvers/base.py:336): a second call adds to the first
    F2 = raytrace(model, srf)
    np.testing.assert_allclose(np.asarray(F2), 2 * g["F_pp"], rtol=1e-9, atol=1e-300)


def test_run_stardis_end_to_end(tmp_path, table_paths):
    """run_stardis(config_fname, tracing_lambdas) with a YAML config: spectrum length == N for the three reference
    test configurations' shapes (test_stardis_full.py:5-27), n_threads semantics, shard == columns of the full run."""
    from stardis_b200 import run_stardis
    from stardis_b200 import units as u
    from stardis_b200.device import default_context

    mod = write_marcs_mod(str(tmp_path / "sun.mod"))
    cfg = tmp_path / "cfg.yml"
    cfg.write_text(f"""
stardis_config_version: 1.0
atom_data: "synthetic:500:3"
input_model:
    type: marcs
    fname: {mod}
    final_atomic_number: 30
opacity:
    file:
        Hminus_bf: {table_paths['Hminus_bf']}
        Hminus_ff: {table_paths['Hminus_ff']}
        H2plus_bf: {table_paths['H2plus_bf']}
    bf:
        H_I: {{}}
    ff:
        H_I: {{}}
    line:
        disable: False
        broadening: [radiation, linear_stark, quadratic_stark, van_der_waals]
no_of_thetas: 10
result_options:
    return_model: true
    return_radiation_field: true
""")
    lam = u.Quantity(np.arange(6560.0, 6570.0, 0.1), u.AA)  # conftest.py:52-56 of the reference
    out = run_stardis(str(cfg), lam)
    assert len(out.spectrum_nu) == len(lam) and len(out.spectrum_lambda) == len(lam)
    assert np.all(np.isfinite(out.spectrum_nu.value)) and np.all(out.spectrum_nu.value > 0)
    assert out.stellar_model.no_of_depth_points == 56 and not hasattr(out, "stellar_plasma")
    F = np.asarray(out.stellar_radiation_field.F_nu)
    np.testing.assert_array_equal(out.spectrum_nu.value, F[-1])
    np.testing.assert_allclose(out.spectrum_lambda.value, F[-1] * out.nus.value / out.lambdas.value)
    # "parallel" config == serial config (test_stardis_full.py:17-27): thread count is irrelevant on the GPU
    out_par = run_stardis(str(cfg), lam, add_config_dict={"n_threads": 4})
    np.testing.assert_array_equal(out_par.spectrum_nu.value, out.spectrum_nu.value)
    with pytest.raises(ValueError):  # stardis/base.py:78-81: 0 is rejected although the schema documents it
        run_stardis(str(cfg), lam, add_config_dict={"n_threads": 0})
    # a nu shard reproduces the same columns bit for bit
    out_sh = run_stardis(str(cfg), lam, shard=(30, 77), add_config_dict={"result_options.return_radiation_field": True})
    np.testing.assert_array_equal(np.asarray(out_sh.stellar_radiation_field.F_nu), F[:, 30:77])
    default_context().synchronize()


# ------------------------------------------------------------------ the other BASELINE.json configurations
def _oracle_total_and_flux(oracle, w, cfg, p0, p1, table_paths, flags=15):
    """CPU oracle on the pixel shard [p0, p1) of a workload: total opacity and F_nu."""
    from stardis_b200 import units as u
    from stardis_b200.constants import H_CGS

    model, plasma, nus = w["model"], w["plasma"], w["nus"]
    T = u.values_of(model.temperatures)
    lt = plasma._line_table.with_masses(model.composition.nuclide_masses).in_range(nus.min(), nus.max())
    lines = {k: getattr(lt, k) for k in ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_upper",
                                         "level_energy_lower", "A_ul", "mass")}
    n_e, n_H = plasma.electron_densities.values, plasma.ion_number_density.loc[1, 0].values
    gam, dws = oracle.calc_broadening(lines, T, n_e, n_H, float(u.cgs_values_of(model.microturbulence)), flags)
    total = oracle.calc_alan_entries(len(T), nus, lt.nu, dws, gam, lt.alpha_line, p0=p0, p1=p1)
    sub = nus[p0:p1]
    nu_cut = (float(plasma.ionization_data.loc[(1, 1)]) - plasma.excitation_energy.values) / H_CGS
    total = total + oracle.alpha_file(sub, T, table_paths["Hminus_bf"], "Hminus_bf", plasma.h_minus_density.values)
    total = total + oracle.alpha_file(sub, T, table_paths["Hminus_ff"], "Hminus_ff", (plasma.ion_number_density.loc[1, 0] * plasma.electron_densities).values)
    total = total + oracle.alpha_bf(sub, nu_cut, np.ones(len(nu_cut)), plasma.level_number_density.values)
    total = total + oracle.alpha_ff(sub, [(1, n_e * plasma.ion_number_density.loc[1, 1].values)], T)
    total = total + oracle.alpha_rayleigh(sub, n_H, plasma.ion_number_density.loc[2, 0].values, plasma.h2_density.values, ["H", "He", "H2"])
    total = total + oracle.alpha_electron(n_e, p1 - p0)
    th, wts = oracle.thetas_and_weights(cfg.no_of_thetas)
    F, _ = oracle.raytrace(T, total, sub, th, wts, dist=model.geometry.dist_to_next_depth_point)
    return total, F


@pytest.mark.parametrize("name", ["solar_full", "astar", "coolgiant_ir"])
def test_full_size_grids_vs_oracle_on_shards(name, table_paths, oracle):
    """BASELINE.json configs 2-4 at their FULL grid sizes (N = 0.55-2.1 M) with a reduced synthetic line list: the GPU
    computes the whole grid (far-field expansion active), the CPU oracle three 256-pixel shards of it (global windows)."""
    from stardis_b200 import units as u
    from stardis_b200.io.config import Configuration, validate_config
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.opacities.opacities_solvers import calc_alphas
    from stardis_b200.radiation_field.radiation_field_solvers import raytrace
    from stardis_b200.synthetic import make_workload

    w = make_workload(name, seed=2, n_lines=2500, strong_fraction=0.01)
    cfg = Configuration(validate_config(dict(
        stardis_config_version=1.0, atom_data="synthetic:2500", input_model=dict(type="marcs", fname="x.mod"),
        opacity=dict(file={"Hminus_bf": table_paths["Hminus_bf"], "Hminus_ff": table_paths["Hminus_ff"]}, bf={"H_I": {}},
                     ff={"H_I": {}}, rayleigh=["H", "He", "H2"],
                     line=dict(broadening=["radiation", "linear_stark", "quadratic_stark", "van_der_waals"])),
        no_of_thetas=w["no_of_thetas"])))
    model, plasma, nus = w["model"], w["plasma"], w["nus"]
    N = len(nus)
    srf = RadiationField(u.Quantity(nus, u.Hz), None, model, cfg.no_of_thetas)
    calc_alphas(plasma, model, srf, cfg.opacity, store_components=False)
    raytrace(model, srf)
    total = np.asarray(srf.opacities.total_alphas)
    F = np.asarray(srf.F_nu)
    assert total.shape == (56, N) and np.isfinite(total).all() and np.isfinite(F).all() and (F[-1] > 0).all()
    for p0 in (0, N // 2 - 77, N - 256):
        ref_total, ref_F = _oracle_total_and_flux(oracle, w, cfg, p0, p0 + 256, table_paths)
        np.testing.assert_allclose(total[:, p0:p0 + 256], ref_total, rtol=RTOL_ALPHA)
        np.testing.assert_allclose(F[:, p0:p0 + 256], ref_F, rtol=RTOL_F, atol=1e-300)
        np.testing.assert_allclose(total[:, p0:p0 + 256], ref_total, rtol=1e-10)  # in fact far tighter


def test_molecular_lines_vs_oracle(table_paths, oracle):
    """include_molecules: second line table, radiation-only gammas of shape (L,1), summed constituent masses
    (opacities_solvers/base.py:444-484, broadening.py:735-821)."""
    from oracle.make_golden_pipeline import CASES, case_inputs
    from stardis_b200 import units as u
    from stardis_b200.plasma.synthetic import attach_synthetic_molecules
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.radiation_field.opacities.opacities_solvers import calc_alphas, calc_molecular_alpha_line_at_nu

    opacity = dict(CASES["bench"])
    opacity["line"] = dict(opacity["line"], include_molecules=True)
    cfg, model, plasma, nus = case_inputs("bench", opacity, table_paths)
    T = u.values_of(model.temperatures)
    attach_synthetic_molecules(plasma, T, 250, nus.min() * 0.9995, nus.max() * 1.0005, seed=4)
    q = u.Quantity(nus, u.Hz)
    srf = RadiationField(q, None, model, 3)
    calc_alphas(plasma, model, srf, cfg.opacity)
    od = srf.opacities.opacities_dict
    assert list(od)[-3:] == ["molecule_alpha_line_at_nu", "molecule_alpha_line_at_nu_gammas", "molecule_alpha_line_at_nu_doppler_widths"]
    ml = plasma.molecule_lines_from_linelist.sort_values("nu")
    ml = ml[ml.nu.between(nus.min(), nus.max())]
    al = plasma.molecule_alpha_line_from_linelist.sort_values("nu")
    al = al[al.nu.between(nus.min(), nus.max())].drop(labels="nu", axis=1).to_numpy()
    ions = plasma.molecule_ion_map.loc[ml.molecule]
    masses = (model.composition.nuclide_masses.loc[ions.Ion1].values + model.composition.nuclide_masses.loc[ions.Ion2].values)
    vmic = float(u.cgs_values_of(model.microturbulence))
    dws = np.array([[oracle.calc_doppler_width(n, t, m, vmic) for t in T] for n, m in zip(ml.nu.values, masses)])
    gam = ml.A_ul.values[:, None]
    ref = oracle.calc_alan_entries(len(T), nus, ml.nu.values, dws, gam, al)
    assert 0 < len(ml) < 250 and np.asarray(od["molecule_alpha_line_at_nu_gammas"]).shape == (len(ml), 1)
    np.testing.assert_allclose(np.asarray(od["molecule_alpha_line_at_nu"]), ref, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(np.asarray(od["molecule_alpha_line_at_nu_doppler_widths"]), dws, rtol=1e-13)
    # total = atomic case + molecular term; atomic gammas still retrievable after the molecular pass
    g = golden("pipeline_golden.npz")
    np.testing.assert_allclose(np.asarray(srf.opacities.total_alphas), g["bench__total"] + ref, rtol=1e-9)
    np.testing.assert_allclose(np.asarray(od["alpha_line_at_nu_gammas"]), g["bench__alpha_line_at_nu_gammas"], rtol=1e-12)
    a, gm, dw = calc_molecular_alpha_line_at_nu(plasma, model, q, cfg.opacity.line)
    np.testing.assert_array_equal(a, np.asarray(od["molecule_alpha_line_at_nu"]))
    cfg.opacity.line.broadening.remove("radiation")  # the reference's latent AttributeError (broadening.py:802-806)
    with pytest.raises(AttributeError):
        calc_molecular_alpha_line_at_nu(plasma, model, q, cfg.opacity.line)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["solar_full", "solar_weak"])
def test_benched_workload_vs_oracle_on_shards(name, oracle):
    """The workload bench.py times, at its FULL size (N = 700 000, L = 300 000 lines, seed 1; 38 % of the (line, depth)
    pairs have whole-grid windows, ~1e5 far-field terms per pixel) and its weak-line twin: GPU on the whole grid through
    the same device step as the bench, CPU oracle on three 256-pixel shards with global windows (a few seconds each)."""
    import argparse

    import bench
    from oracle import ref_leg
    from stardis_b200 import _lib as L
    from stardis_b200.device import default_context

    import torch

    args = argparse.Namespace(workload=name, lines=None)
    w, cfg, _ = bench.build_workload(args)
    ctx = default_context()
    ctx.evict()
    hp = bench.HotPath(ctx, w, cfg, 0, 1, torch)
    hp.step()
    ctx.synchronize()
    total, F = ctx.get(L.BUF_TOTAL), ctx.get(L.BUF_F_NU)
    N = hp.N
    assert total.shape == (56, N) and np.isfinite(total).all() and np.isfinite(F).all() and (F[-1] > 0).all()
    inp = ref_leg.workload_inputs(w, cfg)
    gam, dws = oracle.calc_broadening(inp["lines"], inp["T"], inp["n_e"], inp["n_H"], inp["vmic"], 15)
    th, wts = oracle.thetas_and_weights(inp["n_theta"])
    for p0 in (1024, N // 3 + 11, N - 256):
        a_line = oracle.calc_alan_entries(56, inp["nus"], inp["lines"]["nu"], dws, gam, inp["alpha_line"], p0=p0, p1=p0 + 256)
        ref_total = ref_leg.continuum_total(oracle, inp, inp["nus"][p0:p0 + 256]) + a_line
        ref_F, _ = oracle.raytrace(inp["T"], ref_total, inp["nus"][p0:p0 + 256], th, wts, dist=inp["dist"])
        np.testing.assert_allclose(total[:, p0:p0 + 256], ref_total, rtol=RTOL_ALPHA)
        np.testing.assert_allclose(F[:, p0:p0 + 256], ref_F, rtol=RTOL_F, atol=1e-300)
