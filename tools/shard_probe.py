"""Per-shard K2 time on ONE GPU (emulates the ranks of an N-GPU run one after the other).  Not the bench."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stardis_b200.device import DeviceContext  # noqa: E402
from stardis_b200.distributed import line_balanced_bounds, shard_bounds  # noqa: E402
from stardis_b200.synthetic import make_workload  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 300000
only = int(sys.argv[3]) if len(sys.argv) > 3 else -1
balanced = len(sys.argv) > 4 and sys.argv[4] == "bal"
ctx = DeviceContext(0)
w = make_workload("solar_full", n_lines=n_lines)
p, m = w["plasma"], w["model"]
lt = p.line_table.with_masses(m.composition.nuclide_masses)
ctx.set_atmosphere(w["atmosphere"]["T"], p.electron_densities.values, p.ion_number_density.loc[1, 0].values,
                   w["atmosphere"]["vmic"])
ctx.set_lines(lt.nu, lt.alpha_line, mass=lt.mass, atomic_number=lt.atomic_number, ion_number=lt.ion_number,
              ionization_energy=lt.ionization_energy, level_energy_upper=lt.level_energy_upper,
              level_energy_lower=lt.level_energy_lower, A_ul=lt.A_ul)
N = len(w["nus"])
for r in range(world):
    if only >= 0 and r != only:
        continue
    p0, p1 = line_balanced_bounds(w["nus"], lt.nu, world)[r] if balanced else shard_bounds(N, r, world)
    ctx.set_grid(w["nus"], p0, p1)
    ctx.calc_broadening(15)
    ts = []
    for rep in range(3):
        ctx.timer_start()
        ctx.calc_alpha_line(0)
        ts.append(ctx.timer_stop())
    nl = int(((lt.nu >= min(w["nus"][p0], w["nus"][p1 - 1])) & (lt.nu <= max(w["nus"][p0], w["nus"][p1 - 1]))).sum())
    print(f"rank {r}/{world}: pixels [{p0},{p1}) lines inside {nl}  K2 prep+lines {min(ts):.2f} ms")
