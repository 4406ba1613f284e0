"""CPU legs of bench.py: the reference's own numba-parallel path and the C port, timed on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product package ``stardis_b200``; only ``bench.py``
(``--impl reference`` and the ``cpu_baseline`` leg) and the tests call it.

``numba_sample``  the reference's UNMODIFIED functions (``oracle/ref_shim.py`` executes them from ``/root/reference`` in
                  the build container, from ``oracle/_ref`` on the GPU box): ``calc_gamma`` + ``calc_doppler_width``
                  (broadening.py:550-656, :32-71), ``calc_alan_entries`` (opacities_solvers/base.py:487-592),
                  ``single_theta_trace_parallel`` x n_theta + flux quadrature (radiation_field_solvers/base.py:85-346).
                  The full flagship step would take hours on CPU cores, so one step is a BOUNDED SAMPLE:
                  the line kernel runs on the FULL grid for every k-th line of the nu-sorted list (windows, d_nu and the
                  per-thread scratch slabs are exactly those of the full run), the formal solution on a contiguous block
                  of pixels; the full-step time is t_fixed + k * t_lines + (N / N_block) * t_block.
``port_sample``   ``oracle/stardis_oracle.c`` (OpenMP) on a contiguous nu shard with global windows; also returns the
                  shard's total opacity and flux, which bench.py compares with the GPU result (the parity block).
"""
from __future__ import annotations

import os
import time
import types

import numpy as np


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def workload_inputs(w, cfg):
    """Plain arrays of a ``stardis_b200.synthetic.make_workload`` workload, as both CPU legs consume them."""
    from stardis_b200 import units as u
    from stardis_b200.constants import H_CGS

    model, plasma, nus = w["model"], w["plasma"], w["nus"]
    lt = plasma._line_table
    masses = np.ascontiguousarray(model.composition.nuclide_masses.loc[lt.atomic_number].values, dtype=np.float64)
    lines = dict(nu=lt.nu, atomic_number=lt.atomic_number, ion_number=lt.ion_number, ionization_energy=lt.ionization_energy,
                 level_energy_upper=lt.level_energy_upper, level_energy_lower=lt.level_energy_lower, A_ul=lt.A_ul, mass=masses)
    exc = plasma.excitation_energy.values
    return dict(
        nus=nus, N=len(nus), D=model.no_of_depth_points, T=u.values_of(model.temperatures), lines=lines,
        alpha_line=lt.alpha_line, n_e=plasma.electron_densities.values, n_H=plasma.ion_number_density.loc[1, 0].values,
        n_HII=plasma.ion_number_density.loc[1, 1].values, vmic=float(u.cgs_values_of(model.microturbulence)),
        dist=np.asarray(model.geometry.dist_to_next_depth_point, dtype=np.float64), n_theta=int(cfg.no_of_thetas),
        nu_cut=(float(plasma.ionization_data.loc[(1, 1)]) - exc) / H_CGS, n_level=plasma.level_number_density.values,
        h_minus=plasma.h_minus_density.values, hm_path=cfg.opacity.file.get("Hminus_bf") if hasattr(cfg.opacity.file, "get") else None)


def continuum_total(O, inp, sub):
    """Continuum terms of the bench workloads on the frequencies ``sub`` (numpy restatement, oracle/oracle.py)."""
    total = np.zeros((inp["D"], len(sub)))
    if inp["hm_path"]:
        total = total + O.alpha_file(sub, inp["T"], inp["hm_path"], "Hminus_bf", inp["h_minus"])
    total = total + O.alpha_bf(sub, inp["nu_cut"], np.ones(len(inp["nu_cut"])), inp["n_level"])
    total = total + O.alpha_ff(sub, [(1, inp["n_e"] * inp["n_HII"])], inp["T"])
    return total + O.alpha_electron(inp["n_e"], len(sub))


# ----------------------------------------------------------------------------------------------- C port (oracle/)
def port_sample(w, cfg, target_seconds, threads=None, centre=None, limits=None):
    """One bounded sample with the C/OpenMP port: a contiguous nu shard around pixel ``centre`` (default: the middle of
    the grid) inside ``limits`` (default: the whole grid), sized by a calibration shard so that it takes about
    ``target_seconds``.  Returns a dict with the rate, the description and the shard's (p0, p1, total, F) for the parity
    check."""
    from oracle import oracle as O

    O.build()
    # all host cores this process may use, set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers
    O.set_threads(threads or host_cores())
    cores = O.max_threads()
    inp = workload_inputs(w, cfg)
    nus, N, D, T, lines = inp["nus"], inp["N"], inp["D"], inp["T"], inp["lines"]
    th, wts = O.thetas_and_weights(inp["n_theta"])

    def run(p0, p1):
        t0 = time.perf_counter()
        gam, dws = O.calc_broadening(lines, T, inp["n_e"], inp["n_H"], inp["vmic"], 15)
        t1 = time.perf_counter()
        a_line, evals, _ = O.calc_alan_entries(D, nus, lines["nu"], dws, gam, inp["alpha_line"], p0=p0, p1=p1, with_stats=True)
        total = continuum_total(O, inp, nus[p0:p1]) + a_line
        F, _ = O.raytrace(T, total, nus[p0:p1], th, wts, dist=inp["dist"])
        t2 = time.perf_counter()
        return t2 - t0, evals, t1 - t0, t2 - t1, total, F

    lim0, lim1 = (0, N) if limits is None else (int(limits[0]), int(limits[1]))
    mid = (lim0 + lim1) // 2 if centre is None else int(centre)
    cal = min(64, lim1 - lim0)
    c0 = int(np.clip(mid - cal // 2, lim0, lim1 - cal))
    run(c0, c0 + cal)
    _, _, t_fix, t_var, _, _ = run(c0, c0 + cal)
    width = int(np.clip((target_seconds - t_fix) / max(t_var / cal, 1e-7), cal, lim1 - lim0))
    p0 = int(np.clip(mid - width // 2, lim0, lim1 - width))
    p1 = p0 + width
    t, evals, _, _, total, F = run(p0, p1)
    sample = (f"contiguous nu shard of {p1 - p0} pixels [{p0},{p1}) of the {N}-pixel grid, all {len(lines['nu'])} lines with "
              f"global windows, {evals:.3e} Voigt evaluations, {t:.1f} s on {cores} OpenMP threads (oracle/stardis_oracle.c)")
    return dict(value=(p1 - p0) / t, sample=sample, cores=cores, seconds=t, kind="port", p0=p0, p1=p1, total=total, F=F)


# ----------------------------------------------------------------------------------------------- reference (numba)
def numba_available():
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    from oracle import ref_shim

    return ref_shim.reference_available()


_JIT_WARM = False


def numba_sample(w, cfg, target_seconds, threads=None):
    """One bounded sample of the reference's own numba path (see the module docstring).  Returns a dict like
    ``port_sample`` (without parity arrays), ``kind`` = "reference"."""
    global _JIT_WARM
    cores = int(threads or host_cores())
    os.environ["OMP_NUM_THREADS"] = str(cores)          # torchrun exports 1; numba's omp layer would obey it
    os.environ.setdefault("NUMBA_NUM_THREADS", str(cores))
    import numba

    from oracle import oracle as O
    from oracle.ref_shim import load_reference

    R = load_reference()
    numba.set_num_threads(min(cores, numba.config.NUMBA_NUM_THREADS))
    cores = numba.get_num_threads()
    inp = workload_inputs(w, cfg)
    nus, N, D, T, lines = inp["nus"], inp["N"], inp["D"], inp["T"], inp["lines"]
    L = len(lines["nu"])
    th, wts = O.thetas_and_weights(inp["n_theta"])

    def k1(sel):
        c = {k: v[sel] for k, v in lines.items()}
        gam = R.broadening.calc_gamma(c["atomic_number"][:, None], (c["ion_number"] + 1)[:, None], c["ionization_energy"][:, None],
                                      c["level_energy_upper"][:, None], c["level_energy_lower"][:, None], c["A_ul"][:, None],
                                      inp["n_e"], T, inp["n_H"], True, True, True, True)
        dws = R.broadening.calc_doppler_width(c["nu"][:, None], T, c["mass"][:, None], inp["vmic"])
        return c["nu"], np.ascontiguousarray(gam), np.ascontiguousarray(dws), np.ascontiguousarray(inp["alpha_line"][sel])

    def k2(sel, grid):
        t0 = time.perf_counter()
        lnu, gam, dws, alph = k1(sel)
        t1 = time.perf_counter()
        out = R.opac.calc_alan_entries(D, grid, lnu, dws, gam, alph)
        return t1 - t0, time.perf_counter() - t1, out

    def k4(total, sub):
        t0 = time.perf_counter()
        F = np.zeros_like(total)
        for theta, wt in zip(th, wts):  # raytrace(), radiation_field_solvers/base.py:296-338 (plane-parallel branch)
            I = R.solver.single_theta_trace_parallel(inp["dist"] / np.cos(theta), T.reshape(-1, 1), total, sub,
                                                     R.blackbody.blackbody_flux_at_nu, False)
            F += wt * I
        return time.perf_counter() - t0, F

    if not _JIT_WARM:  # compile every specialisation on a tiny problem (JIT time is not part of any sample)
        k2(np.arange(0, L, max(1, L // 8)), nus[: min(N, 2048)].copy())
        k4(continuum_total(O, inp, nus[:256]) + 1e-12, nus[:256].copy())
        _JIT_WARM = True

    # fixed part of calc_alan_entries on the full grid: per-thread (D, N) scratch slabs zeroed and reduced
    none = np.zeros(0, dtype=np.int64)
    _, t_fixed, _ = k2(none, nus)
    # calibration with 1/2000 of the lines, then the stride for the requested time (2/3 of it for the line kernel)
    k0 = max(1, L // 150)
    t_b0, t_a0, _ = k2(np.arange(k0 // 2, L, k0), nus)
    per_line = max((t_a0 - t_fixed) + t_b0, 1e-6) / max(len(range(k0 // 2, L, k0)), 1)
    n_lines = int(np.clip((0.66 * target_seconds - t_fixed) / per_line, 16, L))
    k = max(1, L // n_lines)
    sel = np.arange(k // 2, L, k)
    t_b, t_a, a_line = k2(sel, nus)
    # formal solution on a contiguous block of pixels in the middle of the grid
    n_block = int(min(N, max(4096, 0.33 * target_seconds * 2.0e4 * cores / 8)))  # ~2e4 nu-points/s for 10 angles on 8 threads
    b0 = (N - n_block) // 2
    sub = nus[b0:b0 + n_block].copy()
    t0 = time.perf_counter()
    total = continuum_total(O, inp, sub) + a_line[:, b0:b0 + n_block]
    t_cont = time.perf_counter() - t0
    t_rt, _ = k4(np.ascontiguousarray(total), sub)
    scale_l, scale_n = L / len(sel), N / n_block
    t_full = t_fixed + (t_b + max(t_a - t_fixed, 0.0)) * scale_l + (t_cont + t_rt) * scale_n
    seconds = t_b + t_a + t_cont + t_rt
    sample = (f"reference numba path ({cores} threads, numba {numba.__version__}): calc_gamma+calc_doppler_width+"
              f"calc_alan_entries on the full {N}-pixel grid for every {k}-th line ({len(sel)} of {L}; {t_b + t_a:.1f} s, of "
              f"which {t_fixed:.1f} s scratch-slab zero/reduce), single_theta_trace_parallel x {inp['n_theta']} on "
              f"{n_block} contiguous pixels ({t_rt:.1f} s), numpy continuum ({t_cont:.2f} s); full step = fixed + "
              f"{scale_l:.1f} x lines + {scale_n:.1f} x block = {t_full:.0f} s")
    return dict(value=N / t_full, sample=sample, cores=cores, seconds=seconds, kind="reference", full_step_seconds=t_full)


def cpu_sample(w, cfg, target_seconds, threads=None, prefer="reference"):
    """The reference's numba path when it can run here, else the C port (says which in ``kind``)."""
    if prefer == "reference" and numba_available():
        return numba_sample(w, cfg, target_seconds, threads)
    return port_sample(w, cfg, target_seconds, threads)
