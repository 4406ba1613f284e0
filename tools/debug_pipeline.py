import sys, os, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import write_table_files
from oracle.make_golden_pipeline import CASES, case_inputs, DEPTH_ROWS
from stardis_b200 import units as u
from stardis_b200.radiation_field import RadiationField
from stardis_b200.radiation_field.opacities.opacities_solvers import calc_alphas
from stardis_b200.radiation_field.radiation_field_solvers import raytrace
g = np.load("tests/golden/pipeline_golden.npz")
tp = write_table_files(tempfile.mkdtemp())
for name in sys.argv[1:] or list(CASES):
    cfg, model, plasma, nus = case_inputs(name, CASES[name], tp)
    srf = RadiationField(u.Quantity(nus, u.Hz), None, model, cfg.no_of_thetas, track_individual_intensities=True)
    total = calc_alphas(plasma, model, srf, cfg.opacity)
    F = raytrace(model, srf)
    for k, v in srf.opacities.opacities_dict.items():
        ref = g[f"{name}__{k}"]
        got = np.asarray(v, dtype=np.float64)
        if ref.ndim == 2 and ref.shape[1] == len(nus):
            got = got[DEPTH_ROWS]
        if got.shape != ref.shape:
            print(name, k, "SHAPE", got.shape, ref.shape); continue
        nanmis = np.isnan(got) != np.isnan(ref)
        with np.errstate(all="ignore"):
            rel = np.abs(got - ref) / np.abs(ref)
        rel = np.where(np.isnan(rel), 0, rel)
        print(name, k, "nan mismatch", nanmis.sum(), "got nan", np.isnan(got).sum(), "ref nan", np.isnan(ref).sum(), "max rel", rel.max() if rel.size else 0)
        if nanmis.any():
            idx = np.argwhere(nanmis)[:5]; print("   at", idx.tolist())
    with np.errstate(all="ignore"):
        print(name, "total max rel", np.nanmax(np.abs(np.asarray(total) - g[f"{name}__total"]) / g[f"{name}__total"]),
              "F max rel", np.nanmax(np.abs(np.asarray(F) - g[f"{name}__F_nu"]) / np.abs(g[f"{name}__F_nu"]).clip(1e-300)))
