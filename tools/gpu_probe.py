"""Quick device probe: FP64 FMA peak, K1/K2 throughput on a slice of the solar workload (not the bench)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stardis_b200 import _lib as L  # noqa: E402
from stardis_b200.device import DeviceContext  # noqa: E402
from stardis_b200.synthetic import make_workload  # noqa: E402

n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = DeviceContext(0)
print("dfma TFLOP/s", ctx.bench_dfma(4096), ctx.bench_dfma(8192))
import ctypes as C
for v in range(3):
    t = C.c_double(); ctx._ck(ctx.lib.sd_bench_fareval(ctx.h, v, 20000, C.byref(t))); print("fareval probe variant", v, "Gevals/s", t.value)
for mode in range(4):
    t = C.c_double(); ctx._ck(ctx.lib.sd_bench_fp64(ctx.h, mode, 4096, C.byref(t))); print("fp64 probe mode", mode, "TFLOP/s", t.value)
x = 10.0 ** np.random.default_rng(0).uniform(0, 30, 1 << 20)
outs = [np.empty_like(x) for _ in range(3)]
ctx._ck(ctx.lib.sd_debug_rcp(ctx.h, x.size, L.ptr(x), *[L.ptr(o) for o in outs]))
ctx.synchronize()
for name, o in zip(("seed", "seed+newton", "seed+cubic"), outs):
    print("rcp", name, "max rel err", np.max(np.abs(o * x - 1.0)))
w = make_workload("solar_full", n_lines=n_lines)
p, m = w["plasma"], w["model"]
lt = p.line_table.with_masses(m.composition.nuclide_masses)
T = w["atmosphere"]["T"]
ctx.set_atmosphere(T, p.electron_densities.values, p.ion_number_density.loc[1, 0].values, w["atmosphere"]["vmic"])
ctx.set_grid(w["nus"])
ctx.set_lines(lt.nu, lt.alpha_line, mass=lt.mass, atomic_number=lt.atomic_number, ion_number=lt.ion_number,
              ionization_energy=lt.ionization_energy, level_energy_upper=lt.level_energy_upper,
              level_energy_lower=lt.level_energy_lower, A_ul=lt.A_ul)
ctx.synchronize()
for rep in range(3):
    ctx.timer_start()
    ctx.calc_broadening(15)
    t1 = ctx.timer_stop()
    ctx.timer_start()
    ctx.calc_alpha_line(0)
    t2 = ctx.timer_stop()
    print(f"rep {rep}: K1 {t1:.3f} ms, K2(prep+lines) {t2:.3f} ms")
a_far = ctx.get(L.BUF_ALPHA_LINE)
ctx.set_farfield(False)
ctx.calc_alpha_line(0)
ctx.timer_start(); ctx.calc_alpha_line(0); t_direct = ctx.timer_stop()
a_dir = ctx.get(L.BUF_ALPHA_LINE)
print(f"direct mode K2 {t_direct:.3f} ms; far-field vs direct max rel dev {np.max(np.abs(a_far - a_dir) / a_dir):.3e}")
ctx.set_farfield(True)
ctx.calc_alpha_line(0)
ctx.set_line_stats(True)
ctx.calc_alpha_line(0)
st = ctx.line_stats()
ctx.set_line_stats(False)
print("stats", st)
ctx.timer_start()
ctx.calc_alpha_line(0)
t2 = ctx.timer_stop()
print(f"K2 lines only {t2:.3f} ms -> {st['evals'] / t2 / 1e6:.1f} Gevals/s")
a = ctx.get(L.BUF_ALPHA_LINE)
print("alpha_line finite", np.isfinite(a).all(), a.min(), a.max())
