"""CPU legs of bench.py (oracle/ref_leg.py): the C port sample and the reference's own numba path staged in oracle/_ref."""
import argparse

import numpy as np
import pytest


@pytest.fixture(scope="module")
def small_workload():
    import bench

    return bench.build_workload(argparse.Namespace(workload="sim100aa", lines=300))


def test_port_sample_is_the_oracle_on_a_shard(small_workload, oracle):
    from oracle import ref_leg

    w, cfg, _ = small_workload
    r = ref_leg.port_sample(w, cfg, 0.5, limits=(2000, 6000))
    assert r["kind"] == "port" and 2000 <= r["p0"] < r["p1"] <= 6000 and r["value"] > 0
    inp = ref_leg.workload_inputs(w, cfg)
    gam, dws = oracle.calc_broadening(inp["lines"], inp["T"], inp["n_e"], inp["n_H"], inp["vmic"], 15)
    full = oracle.calc_alan_entries(inp["D"], inp["nus"], inp["lines"]["nu"], dws, gam, inp["alpha_line"])
    total = ref_leg.continuum_total(oracle, inp, inp["nus"]) + full
    np.testing.assert_allclose(r["total"], total[:, r["p0"]:r["p1"]], rtol=1e-13)
    assert r["F"].shape == r["total"].shape and (r["F"][-1] > 0).all()


def test_reference_numba_leg_runs_the_staged_modules(small_workload, oracle):
    """The reference's unmodified numba functions (oracle/_ref or /root/reference through oracle/ref_shim.py) agree with
    the C port on the same inputs, and the bounded-sample extrapolation returns a rate."""
    from oracle import ref_leg, ref_shim, stage_ref

    stage_ref.stage()
    if not ref_leg.numba_available():
        pytest.skip("numba or the staged reference modules are not available")
    w, cfg, _ = small_workload
    inp = ref_leg.workload_inputs(w, cfg)
    R = ref_shim.load_reference()
    lines = inp["lines"]
    gam = R.broadening.calc_gamma(lines["atomic_number"][:, None], (lines["ion_number"] + 1)[:, None],
                                  lines["ionization_energy"][:, None], lines["level_energy_upper"][:, None],
                                  lines["level_energy_lower"][:, None], lines["A_ul"][:, None], inp["n_e"], inp["T"], inp["n_H"],
                                  True, True, True, True)
    dws = R.broadening.calc_doppler_width(lines["nu"][:, None], inp["T"], lines["mass"][:, None], inp["vmic"])
    ref = R.opac.calc_alan_entries(inp["D"], inp["nus"], lines["nu"], dws, gam, inp["alpha_line"])
    g2, d2 = oracle.calc_broadening(lines, inp["T"], inp["n_e"], inp["n_H"], inp["vmic"], 15)
    port = oracle.calc_alan_entries(inp["D"], inp["nus"], lines["nu"], d2, g2, inp["alpha_line"])
    np.testing.assert_allclose(port, ref, rtol=1e-10, atol=1e-300)
    r = ref_leg.numba_sample(w, cfg, 1.0)
    assert r["kind"] == "reference" and r["value"] > 0 and r["cores"] >= 1 and "every" in r["sample"]
