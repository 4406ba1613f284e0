"""CPU-only tests: host logic of the drop-in API, the C-ABI library surface (load + symbols, no compute), the shared
scalar arithmetic compiled for the host, and the multi-rank gather logic on gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden
from helpers import write_marcs_mod, write_table_files


# ------------------------------------------------------------------ C ABI surface
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "stardis_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_library():
    from stardis_b200.build import build

    return build()


def test_library_exports_every_declared_symbol(built_library):
    lib = C.CDLL(built_library)  # dlopen works without a GPU (cudart is linked statically)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/stardis_b200.h but not exported"
    from stardis_b200 import _lib

    assert set(_lib.SIGNATURES) == set(names)  # the ctypes binding covers exactly the declared ABI
    _lib.load()
    lib.sd_version.restype = C.c_char_p
    assert b"sm_100a" in lib.sd_version()


def test_library_contains_sm100a_code(built_library):
    out = subprocess.run(["cuobjdump", "--list-elf", built_library], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from stardis_b200._lib import StardisB200Error
    from stardis_b200.device import DeviceContext
    from stardis_b200.radiation_field.opacities.opacities_solvers import voigt_profile

    with pytest.raises(StardisB200Error):
        DeviceContext(0)
    with pytest.raises(StardisB200Error):  # no silent CPU fallback behind the public functions either
        voigt_profile(0.0, 1.0, 0.0)


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "stardis_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "stardis_oracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


# ------------------------------------------------------------------ shared scalar arithmetic on the host
@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "libhm.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", so,
                    os.path.join(ROOT, "tests", "host_math_harness.cpp")], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_kernel_arithmetic_vs_reference_golden(host_math):
    """stardis_b200/csrc/sd_math.cuh (the formulas every kernel uses) compiled for the host, against the reference."""
    g = golden("kernels_golden.npz")
    z = g["fad_z"]
    x, y = np.ascontiguousarray(z.real), np.ascontiguousarray(z.imag)
    n = x.size
    re_, wr, wi = np.empty(n), np.empty(n), np.empty(n)
    reg = np.empty(n, dtype=np.int32)
    host_math.hm_humlicek(C.c_long(n), _p(x), _p(y), _p(re_), _p(wr), _p(wi), reg.ctypes.data_as(C.POINTER(C.c_int)))
    assert set(np.unique(reg)) == {0, 1, 2, 3}
    np.testing.assert_allclose(re_, g["fad_w"].real, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(wr, g["fad_w"].real, rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(wi, g["fad_w"].imag, rtol=1e-12, atol=1e-18)
    phi = np.empty(g["vp_dnu"].size)
    host_math.hm_voigt(C.c_long(phi.size), _p(g["vp_dnu"]), _p(g["vp_dw"]), _p(g["vp_gamma"]), _p(phi))
    np.testing.assert_allclose(phi, g["vp_phi"], rtol=1e-13)
    m = g["b_nu"].size
    zeff = g["b_zeff"].astype(np.float64)
    ls, qs, vw = np.empty(m), np.empty(m), np.empty(m)
    host_math.hm_broadening(C.c_long(m), _p(zeff), _p(g["b_nu"]), _p(g["b_nl"]), _p(g["b_ne"]), _p(g["b_T"]), _p(g["b_nH"]),
                            _p(ls), _p(qs), _p(vw))
    np.testing.assert_allclose(ls, g["b_linear_stark"], rtol=1e-13)
    np.testing.assert_allclose(qs, g["b_quadratic_stark"], rtol=1e-13)
    np.testing.assert_allclose(vw, g["b_van_der_waals"], rtol=1e-13)
    neff, dw = np.empty(m), np.empty(m)
    host_math.hm_neff_doppler(C.c_long(m), _p(zeff), _p(g["b_eion"]), _p(g["b_elev"]), _p(g["b_nuline"]), _p(g["b_T"]),
                              _p(g["b_mass"]), C.c_double(1.3e5), _p(neff), _p(dw))
    np.testing.assert_allclose(neff, g["b_neff"], rtol=1e-14, equal_nan=True)
    np.testing.assert_allclose(dw, g["b_doppler"], rtol=1e-14)


def test_window_rule_vs_oracle(host_math, oracle):
    """sdm::line_window (used by k_build_records) == the oracle's window, incl. the NaN / overflow quirks."""
    g = golden("alan_golden.npz")
    nus, ln, dws, gam, al = (g[f"b_{k}"] for k in ("nus", "line_nus", "dws", "gammas", "alphas"))
    lo_ref, hi_ref = oracle.line_windows(nus, ln, dws, gam, al)
    idx = np.array([(nus >= v).sum() for v in ln], dtype=np.int64)
    L, D = dws.shape
    lo, hi = np.empty(L * D, dtype=np.int64), np.empty(L * D, dtype=np.int64)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))  # noqa: E731
    idx_ld = np.ascontiguousarray(np.repeat(idx, D))
    host_math.hm_window(C.c_long(L * D), ip(idx_ld), C.c_longlong(nus.size), _p(np.ascontiguousarray(gam).ravel()),
                        _p(np.ascontiguousarray(dws).ravel()), _p(np.ascontiguousarray(al).ravel()),
                        C.c_double(oracle.d_nu(nus)), ip(lo), ip(hi))
    assert np.array_equal(lo.reshape(L, D), lo_ref) and np.array_equal(hi.reshape(L, D), hi_ref)
    assert (hi_ref[9, 1] - lo_ref[9, 1]) == 0 and (hi_ref[11, 0] - lo_ref[11, 0]) == 0  # inf / 1e40 alpha: empty window
    assert (hi_ref[5] - lo_ref[5] <= 20).all()  # NaN gamma: forced to the 10-pixel half-width


# ------------------------------------------------------------------ configuration
def _raw_config(**over):
    raw = dict(stardis_config_version=1.0, atom_data="synthetic:10", input_model=dict(type="marcs", fname="sun.mod"),
               opacity=dict(line={}))
    raw.update(over)
    return raw


def test_config_defaults_follow_the_reference_schemas():
    from stardis_b200.io.config import Configuration, validate_config

    cfg = Configuration(validate_config(_raw_config()))
    assert cfg.n_threads == 1 and cfg.no_of_thetas == 10
    assert cfg.input_model.final_atomic_number == 92 and cfg.input_model.gzipped is False
    assert cfg.opacity.file == {} and cfg.opacity.bf == {} and cfg.opacity.rayleigh == []
    assert cfg.opacity.disable_electron_scattering is False
    line = cfg.opacity.line
    assert line.disable is False and line.broadening == [] and line.include_molecules is False
    assert line.vald_linelist.use_linelist is False and line.vald_linelist.use_vald_broadening is True
    assert cfg.result_options.return_radiation_field is False
    # "x" in broadening membership tests (broadening.py:688-691)
    cfg2 = Configuration(validate_config(_raw_config(opacity=dict(line=dict(broadening=["radiation", "van_der_waals"])))))
    assert "radiation" in cfg2.opacity.line.broadening and "linear_stark" not in cfg2.opacity.line.broadening
    # the reference's own test configs say "model:" instead of "input_model:"
    raw = _raw_config()
    raw["model"] = raw.pop("input_model")
    assert Configuration(validate_config(raw)).input_model.type == "marcs"
    cfg.set_config_item("opacity.line.disable", True)
    assert cfg.opacity.line.disable is True


@pytest.mark.parametrize("bad", [
    dict(stardis_config_version=2.0), dict(opacity=dict(file={"Hminus": "x.dat"})), dict(opacity=dict(rayleigh=["Fe"])),
    dict(opacity=dict(line=dict(broadening=["doppler"]))), dict(opacity=dict(unknown=1)),
    dict(input_model=dict(type="phoenix", fname="x")), dict(result_options=dict(return_everything=True))])
def test_config_rejects_invalid_input(bad):
    from stardis_b200.io.config import validate_config

    with pytest.raises(ValueError):
        validate_config(_raw_config(**bad))


def test_set_num_threads_semantics():
    from stardis_b200 import set_num_threads

    for ok in (1, 4, -99):
        set_num_threads(ok)
    for bad in (0, -1):  # stardis/base.py:78-81 (0 is documented by the schema but rejected by the code)
        with pytest.raises(ValueError):
            set_num_threads(bad)


# ------------------------------------------------------------------ model IO, units
def test_marcs_reader_and_stellar_model(tmp_path):
    from stardis_b200.io.model.marcs import read_marcs_model
    from stardis_b200.synthetic import load_atmosphere

    m = read_marcs_model(write_marcs_mod(str(tmp_path / "sun.mod")))
    atm = load_atmosphere("sun")
    assert not m.spherical and m.data.shape[0] == 56
    assert m.metadata["teff"].value == 5777.0 and m.metadata["microturbulence"].cgs.value == 1.0e5
    sm = m.to_stellar_model(final_atomic_number=30)
    np.testing.assert_allclose(sm.temperatures.value, atm["T"])           # deepest point first (marcs.py:203-205)
    np.testing.assert_allclose(sm.geometry.r.value, atm["r"], rtol=1e-3)  # 4 significant digits in the file
    assert sm.no_of_depth_points == 56 and sm.geometry.dist_to_next_depth_point.shape == (55,)
    assert (sm.geometry.dist_to_next_depth_point > 0).all()
    assert abs(sm.composition.nuclide_masses.loc[26] / 1.66053906660e-24 - 55.845) < 1e-9
    np.testing.assert_allclose(sm.composition.elemental_mass_fraction[0].sum(), 1.0)


def test_units():
    from stardis_b200 import units as u

    lam = u.Quantity(np.array([3000.0, 10000.0]), u.AA)
    nu = lam.to(u.Hz, u.spectral())
    np.testing.assert_allclose(nu.value, 2.99792458e18 / lam.value)
    np.testing.assert_allclose(nu.to(u.AA, u.spectral()).value, lam.value)
    np.testing.assert_allclose(u.to_hz(lam).value, nu.value)
    assert u.Quantity(1.5, u.km_s).cgs.value == 1.5e5
    assert float(u.Quantity(np.array([5.0, 6.0]), u.K)[1].value) == 6.0
    with pytest.raises(ValueError):
        lam.to(u.Hz)
    with pytest.raises(TypeError):
        u.to_hz(np.array([1.0]))


# ------------------------------------------------------------------ line tables
def test_columnar_lines_from_pandas_plasma_matches_native_table():
    from stardis_b200.plasma.columnar import ColumnarLines
    from stardis_b200.synthetic import load_atmosphere
    from stardis_b200.plasma.synthetic import create_synthetic_plasma

    atm = load_atmosphere("sun")
    for vald in (False, True):
        plasma = create_synthetic_plasma(atm, 500, 4.5e14, 4.6e14, seed=3, vald=vald)
        native = plasma._line_table
        native.level_energy_upper[::50] = native.ionization_energy[::50] * 1.01  # some auto-ionising lines
        table = ColumnarLines.from_plasma(plasma, use_vald=vald)
        for k in ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_lower", "level_energy_upper",
                  "A_ul", "alpha_line"):
            np.testing.assert_array_equal(getattr(table, k), getattr(native, k), err_msg=k)
        if vald:
            np.testing.assert_array_equal(table.waals, native.waals)
        assert (np.diff(table.nu) >= 0).all()
        sub = table.in_range(table.nu[10], table.nu[20])  # pandas between() is inclusive on both ends
        assert len(sub) == 11 and sub.nu[0] == table.nu[10] and sub.nu[-1] == table.nu[20]
        kept = table.without_autoionizing()
        assert len(kept) == len(table) - 10 and not (kept.level_energy_upper > kept.ionization_energy).any()
        assert table.without_autoionizing() is kept  # cached


def test_cross_section_tables_and_delaunay_split(tmp_path, oracle):
    """Table parsing + the per-cell diagonal flags reproduce scipy's LinearNDInterpolator (what the reference calls)
    when evaluated with the kernel's triangle formula (restated here in numpy)."""
    from stardis_b200.radiation_field.opacities.opacities_solvers.util import read_table, table_descriptor

    paths = write_table_files(str(tmp_path))
    rng = np.random.default_rng(0)
    T = np.array([3500.0, 5040.0, 7777.0, 9900.0, 2000.0, 30000.0])
    lam = np.concatenate([rng.uniform(900, 40000, 400), [1823.0, 151890.0, 5000.0, 9113.0]])
    for src in ("Hminus_ff", "H2plus_bf"):
        t = read_table(paths[src], src)
        ref = oracle.sigma_file(lam, T, paths[src], src)
        ycoord = 5040.0 / T if src == "Hminus_ff" else T
        scale = 1e-26 * 1.380649e-16 * T if src == "Hminus_ff" else np.full_like(T, 1e-18)
        got = np.zeros_like(ref)
        xs, ys, v, dg = t["x"], t["y"], t["values"], t["diag"]
        for d, yq in enumerate(ycoord):
            for i, xq in enumerate(lam):
                if not (xs[0] <= xq <= xs[-1] and ys[0] <= yq <= ys[-1]):
                    continue
                ix = min(np.searchsorted(xs, xq, side="right") - 1, len(xs) - 2)
                jy = min(np.searchsorted(ys, yq, side="right") - 1, len(ys) - 2)
                fx = (xq - xs[ix]) / (xs[ix + 1] - xs[ix])
                fy = (yq - ys[jy]) / (ys[jy + 1] - ys[jy])
                v00, v01, v10, v11 = v[ix, jy], v[ix, jy + 1], v[ix + 1, jy], v[ix + 1, jy + 1]
                if dg[ix, jy] == 0:
                    val = v00 + fx * (v10 - v00) + fy * (v11 - v10) if fx >= fy else v00 + fy * (v01 - v00) + fx * (v11 - v01)
                else:
                    val = (v00 + fx * (v10 - v00) + fy * (v01 - v00) if fx + fy <= 1 else
                           v11 + (1 - fx) * (v01 - v11) + (1 - fy) * (v10 - v11))
                got[d, i] = val * scale[d]
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-300)
        assert (ref == 0).any() and (ref > 0).any()  # both inside and outside the table hull were probed
    t1 = table_descriptor(paths["Hminus_bf"], "Hminus_bf", T, np.ones_like(T))
    assert t1["kind"] == 1 and t1["depth_y"] is None
    with pytest.raises(ValueError):
        read_table(paths["Hminus_ff"], "Heminus_ff")


def test_continuum_descriptors_vs_oracle(oracle):
    """bf prefix sums / ff coefficients (host side of K3) reproduce the oracle's per-level evaluation."""
    from stardis_b200.io.config import Configuration
    from stardis_b200.radiation_field.opacities.opacities_solvers import base as ob
    from stardis_b200.synthetic import load_atmosphere, stellar_model_from_atmosphere
    from stardis_b200.plasma.synthetic import create_synthetic_plasma

    atm = load_atmosphere("sun")
    model = stellar_model_from_atmosphere(atm)
    plasma = create_synthetic_plasma(atm, 10, 4e14, 5e14)
    nus = 2.99792458e18 / np.linspace(800.0, 30000.0, 700)
    cut, prefix = ob.bf_descriptor(plasma, Configuration({"H_I": {}}))
    k = np.searchsorted(cut, nus, side="right")
    mine = prefix[k].T * nus ** -3.0
    exc = plasma.excitation_energy.values
    nu_cut = (float(plasma.ionization_data.loc[(1, 1)]) - exc) / 6.62607015e-27
    ref = oracle.alpha_bf(nus, nu_cut, np.ones(len(nu_cut)), plasma.level_number_density.values)
    np.testing.assert_allclose(mine, ref, rtol=1e-13)
    assert (mine[:, -1] < mine[:, 0]).all() and len(np.unique(k)) > 5
    coef = ob.ff_descriptor(plasma, model, Configuration({"H_I": {}}))
    ref = oracle.alpha_ff(nus, [(1, plasma.electron_densities.values * plasma.ion_number_density.loc[1, 1].values)], atm["T"])
    np.testing.assert_allclose(coef[:, None] * nus ** -3.0, ref, rtol=1e-14)
    assert ob.bf_descriptor(plasma, Configuration({})) == (None, None) and ob.ff_descriptor(plasma, model, Configuration({})) is None
    n, z, ion = ob.get_number_density(plasma, "H_I_ff")
    assert (z, ion) == (1, 1)
    n, z, ion = ob.get_number_density(plasma, "H_I_bf")
    assert (z, ion) == (1, 0)


# ------------------------------------------------------------------ containers
class _FakeCtx:
    def __init__(self):
        self.fetches = 0

    def get(self, which, shape=None):
        self.fetches += 1
        return np.full(shape, float(which))

    def get_row(self, which, row):
        return np.full(4, 100.0 + which)


def test_device_array_is_lazy_and_array_like():
    from stardis_b200.device_array import DeviceArray

    ctx = _FakeCtx()
    a = DeviceArray(ctx, 6, (3, 4))
    assert a.shape == (3, 4) and a.ndim == 2 and len(a) == 3 and ctx.fetches == 0
    np.testing.assert_array_equal(a[-1], np.full(4, 106.0))  # row access does not fetch the whole array
    assert ctx.fetches == 0
    assert (a + 1).sum() == 3 * 4 * 7 and ctx.fetches == 1
    assert np.asarray(a).mean() == 6 and a.T.shape == (4, 3) and ctx.fetches == 1  # cached after first use
    assert "host" in repr(a)


def test_opacities_container_total_follows_the_reference():
    from stardis_b200.radiation_field.opacities import Opacities

    class M:
        no_of_depth_points = 2

    op = Opacities(np.zeros(3), M())
    assert op.total_alphas.shape == (2, 3)
    op.opacities_dict.update(alpha_a=np.ones((2, 3)), alpha_b=2 * np.ones((2, 3)), alpha_line_at_nu_gammas=np.full((5, 2), 9.0),
                             alpha_line_at_nu_doppler_widths=np.full((5, 2), 9.0), alpha_electron=0)
    np.testing.assert_array_equal(op.calc_total_alphas(), 3 * np.ones((2, 3)))
    np.testing.assert_array_equal(op.calc_total_alphas(), 6 * np.ones((2, 3)))  # accumulates (opacities/base.py:27)


def test_radiation_field_quadrature_and_output_layout():
    from stardis_b200 import STARDISOutput
    from stardis_b200 import units as u
    from stardis_b200.io.config import Configuration
    from stardis_b200.radiation_field import RadiationField
    from stardis_b200.synthetic import load_atmosphere, stellar_model_from_atmosphere

    model = stellar_model_from_atmosphere(load_atmosphere("sun"))
    lam = u.Quantity(np.array([5000.0, 5001.0, 5002.0]), u.AA)
    nus = u.to_hz(lam)
    srf = RadiationField(nus, None, model, 3, track_individual_intensities=True)
    # radiation_field/base.py:60-63 (values from the reference)
    np.testing.assert_allclose(srf.thetas, [0.3980998287767066, 0.7853981633974483, 1.17269649801819], rtol=1e-15)
    np.testing.assert_allclose(srf.I_nus_weights, [0.872664625997165, 1.3962634015954636, 0.872664625997165], rtol=1e-15)
    assert srf.F_nu.shape == (56, 3) and not srf.F_nu.any() and srf.I_nus.shape == (56, 3, 3)
    srf.F_nu = np.arange(56 * 3, dtype=float).reshape(56, 3)
    ro = Configuration(dict(return_model=True, return_plasma=False, return_radiation_field=False))
    out = STARDISOutput(ro, model, "plasma", srf)
    assert hasattr(out, "stellar_model") and not hasattr(out, "stellar_plasma") and not hasattr(out, "stellar_radiation_field")
    np.testing.assert_array_equal(out.spectrum_nu.value, srf.F_nu[-1])
    np.testing.assert_allclose(out.lambdas.value, lam.value)
    np.testing.assert_allclose(out.spectrum_lambda.value, srf.F_nu[-1] * nus.value / lam.value)
    with pytest.raises(AttributeError):
        RadiationField(nus, None, model, 3).I_nus


def test_spherical_ray_geometry_vs_oracle(oracle):
    from stardis_b200.radiation_field.radiation_field_solvers.base import calculate_spherical_ray

    g = golden("raytrace_golden.npz")
    th, _ = oracle.thetas_and_weights(4)
    np.testing.assert_allclose(calculate_spherical_ray(th, g["r_sph"]), g["sph_ray"], rtol=1e-13, atol=1e-300)


# ------------------------------------------------------------------ multi-rank gather (gloo, world_size 2)
_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from stardis_b200.distributed import (shard_bounds, allgather_spectrum, allgather_columns, line_balanced_bounds,
                                      upload_rows_striped, stripe_rows)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
N, D = 1001, 5
full = np.arange(N, dtype=np.float64) ** 1.5
cols = np.arange(D * N, dtype=np.float64).reshape(D, N)
p0, p1 = shard_bounds(N, rank, world)
spec = allgather_spectrum(full[p0:p1].copy(), (p0, p1), N)
mat = allgather_columns(np.ascontiguousarray(cols[:, p0:p1]), (p0, p1), N)
ok = np.array_equal(spec, full) and np.array_equal(mat, cols)
try:
    allgather_spectrum(full[p0:p1 - 1].copy(), (p0, p1 - 1), N)
    ok = False
except ValueError:
    pass
# unequal, cost-balanced ranges: twice as many lines per pixel in the first quarter of the grid
nus = np.linspace(9.0e14, 3.0e14, N)
line_nus = np.concatenate([np.linspace(8.9e14, 7.5e14, 600), np.linspace(7.4e14, 3.1e14, 300)])
bounds = line_balanced_bounds(nus, line_nus, world, align=16)
q0, q1 = bounds[rank]
ok = ok and (bounds[0][1] - bounds[0][0]) < (bounds[1][1] - bounds[1][0])
spec = allgather_spectrum(full[q0:q1].copy(), (q0, q1), N, bounds=bounds)
mat = allgather_columns(np.ascontiguousarray(cols[:, q0:q1]), (q0, q1), N, bounds=bounds)
ok = ok and np.array_equal(spec, full) and np.array_equal(mat, cols)
try:
    allgather_spectrum(full[q0:q1].copy(), (q0, q1), N, bounds=[(0, 10), (20, N)])
    ok = False
except ValueError:
    pass
# striped upload of a table every rank holds: each rank contributes its row block, everyone ends up with all rows
table = np.arange(37 * 3, dtype=np.float64).reshape(37, 3) ** 1.1
got = upload_rows_striped(table, "cpu")
ok = ok and tuple(got.shape) == (37, 3) and np.array_equal(got.numpy(), table)
ok = ok and stripe_rows(37, 0, 2) == (0, 19, 19) and stripe_rows(37, 1, 2) == (19, 37, 19) and stripe_rows(1, 1, 2) == (1, 1, 1)
got1 = upload_rows_striped(table[:1], "cpu")   # fewer rows than ranks: one rank contributes nothing
ok = ok and np.array_equal(got1.numpy(), table[:1])
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 3)
"""


def test_nu_shards_gather_to_the_full_spectrum_gloo_world2(tmp_path):
    from stardis_b200.distributed import all_shards, shard_bounds

    for n, w in ((700000, 8), (1001, 2), (7, 3)):
        b = all_shards(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(q - p for p, q in b) - min(q - p for p, q in b) <= 1
        assert shard_bounds(n, w - 1, w) == b[-1]
    from stardis_b200.distributed import line_balanced_bounds

    lam = np.arange(3000, 10000, 0.01)
    nus = 2.99792458e18 / lam
    line_nus = np.random.default_rng(5).uniform(nus.min(), nus.max(), 50000)
    idx = len(nus) - np.searchsorted(nus[::-1], line_nus)
    for w in (1, 2, 8):
        b = line_balanced_bounds(nus, line_nus, w, line_weight=4.0, align=512)
        assert b[0][0] == 0 and b[-1][1] == len(nus) and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(p % 512 == 0 for p, _ in b)
        cost = [(q - p) + 4.0 * ((idx >= p) & (idx < q)).sum() for p, q in b]
        assert max(cost) - min(cost) <= 2 * (512 + 4.0 * 512 * 50000 / len(nus) * 12)  # within about one cut of equal
    assert line_balanced_bounds(nus, np.zeros(0), 3) == [(0, 233472), (233472, 466944), (466944, 700000)]
    assert all(q > p for p, q in line_balanced_bounds(nus[:100], line_nus, 4))
    script = tmp_path / "worker.py"
    port = 29500 + os.getpid() % 2000
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(os.environ, RANK=str(r), WORLD_SIZE="2"))
             for r in range(2)]
    assert [p.wait(timeout=240) for p in procs] == [0, 0]


# ------------------------------------------------------------------ VALD line strengths: host preparation
@pytest.mark.parametrize("kind", ["long", "short"])
def test_vald_linelist_preparation_matches_the_reference_line_table(kind):
    from stardis_b200.plasma.alpha_line_vald import prepare_vald_linelist

    g = golden("plasma_golden.npz")
    ll = {k[3:]: g[k] for k in g.files if k.startswith("ll_")}
    lines = prepare_vald_linelist(ll, g["ions"], g["ionization_index"], g["ionization_energy"], int(g["max_atomic_number"]),
                                  shortlist=(kind == "short"))
    assert len(lines) == g[f"{kind}_alpha"].shape[0]
    for col in ("atomic_number", "ion_number", "nu", "level_energy_lower", "level_energy_upper", "A_ul", "ionization_energy"):
        np.testing.assert_allclose(getattr(lines, col), g[f"{kind}_lines_{col}"], rtol=1e-14, err_msg=col)
    assert (lines.g_lo is None) == (kind == "short")
    with pytest.raises(ValueError):
        prepare_vald_linelist(ll, g["ions"][:3], g["ionization_index"], g["ionization_energy"], 28)
