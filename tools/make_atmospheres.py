"""Extract the depth structure of the two MARCS models that ship with the reference into
stardis_b200/data/atmospheres.npz (build container only; the GPU box has no /root/reference).

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
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "stardis_b200", "data", "atmospheres.npz")
np.savez_compressed(dst, **out)
print(dst, os.path.getsize(dst))
