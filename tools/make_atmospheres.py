"""Extract the depth structure of the two MARCS models that ship with the reference into
benchdata/atmospheres.npz (build container only; the GPU box has no /root/reference).

Only the numeric structure columns are stored (depth, T, Pe, Pg, density, microturbulence, Teff); they feed the
synthetic workloads of bench.py and the tests.  Parsed with the product's own MARCS reader.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stardis_b200.io.model.marcs import read_marcs_model  # noqa: E402

REF = os.environ.get("STARDIS_REFERENCE_ROOT", "/root/reference")
SRC = {"sun": "docs/quickstart/sun.mod", "cool": "stardis/io/model/tests/data/marcs_test.mod.gz"}
out = {}
for name, rel in SRC.items():
    m = read_marcs_model(os.path.join(REF, rel))
    d = m.data
    for col in ("depth", "t", "pe", "pg", "density"):
        out[f"{name}_{col}"] = d[col].values[::-1].astype(np.float64).copy()  # deepest -> surface
    out[f"{name}_vmic_kms"] = float(m.metadata["microturbulence"].value)
    out[f"{name}_teff"] = float(m.metadata["teff"].value)
    out[f"{name}_logA"] = m.log_abundances
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchdata", "atmospheres.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst))

# ---- numeric content of the three cross-section tables the reference ships in stardis/data/ (Wishart 1979 H- bf,
# Bell & Berrington 1987 H- ff, Stancil 1994 H2+ bf) -> benchdata/cross_sections.npz
from stardis_b200.radiation_field.opacities.opacities_solvers.util import read_table  # noqa: E402

TABLES = {"Hminus_bf": "h_minus_bf_W1979.dat", "Hminus_ff": "h_minus_ff_B1987.dat", "H2plus_bf": "h2_plus_bf_S1994.dat"}
cs = {}
for src, fn in TABLES.items():
    t = read_table(os.path.join(REF, "stardis", "data", fn), src)
    cs[f"{src}_x"] = t["x"] / (10.0 if src == "H2plus_bf" else 1.0)  # keep the file's own wavelength unit (nm for H2+)
    cs[f"{src}_values"] = t["values"]
    if t["y"] is not None:
        cs[f"{src}_y"] = t["y"]
dst = os.path.join(os.path.dirname(dst), "cross_sections.npz")
np.savez_compressed(dst, **cs)
print(dst, os.path.getsize(dst))
