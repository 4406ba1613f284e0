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
