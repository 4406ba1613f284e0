"""Time K2 on a narrow-lines-only workload (every window is the 20-pixel minimum): the line-core (exact) path."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stardis_b200 import _lib as L
from stardis_b200.device import DeviceContext
from stardis_b200.synthetic import make_workload
n_lines = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
la = (float(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (-6.0, -5.0)
ctx = DeviceContext(0)
w = make_workload("solar_full", n_lines=n_lines, strong_fraction=0.0, log_alpha=la)
p, m = w["plasma"], w["model"]
lt = p.line_table.with_masses(m.composition.nuclide_masses)
ctx.set_atmosphere(w["atmosphere"]["T"], p.electron_densities.values, p.ion_number_density.loc[1, 0].values, w["atmosphere"]["vmic"])
ctx.set_grid(w["nus"])
ctx.set_lines(lt.nu, lt.alpha_line, mass=lt.mass, atomic_number=lt.atomic_number, ion_number=lt.ion_number,
              ionization_energy=lt.ionization_energy, level_energy_upper=lt.level_energy_upper,
              level_energy_lower=lt.level_energy_lower, A_ul=lt.A_ul)
ctx.calc_broadening(15)
ctx.calc_alpha_line(0)
ctx.synchronize()
for rep in range(2):
    ctx.timer_start(); ctx.lib.sd_calc_alpha_line(ctx.h, 0); t = ctx.timer_stop()
ctx.set_line_stats(True); ctx.calc_alpha_line(0); st = ctx.line_stats(); ctx.set_line_stats(False)
print(f"log_alpha {la}: K2 {t:.3f} ms, evals {st['evals']:.3e}, regions {st['region_evals']}, pairs {st['pairs']}, wide {st['wide_pairs']}")
