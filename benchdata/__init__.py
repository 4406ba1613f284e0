"""Bench / test fixtures, NOT part of the product package: MARCS structure columns (atmospheres.npz) and the three
continuum cross-section tables (cross_sections.npz), extracted as plain numbers from the model and table files the
reference ships (tools/make_atmospheres.py; published physics tables / model data).  ``stardis_b200.synthetic`` builds the
bench workloads from them; the product path itself reads whatever files its configuration names."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def write_cross_section_files(dirpath):
    """Write the cross-section tables as text files in the formats ``opacity.file`` expects (the reference's
    stardis/data/*.dat layouts) and return {opacity_source: path}."""
    t = np.load(os.path.join(HERE, "cross_sections.npz"))
    os.makedirs(dirpath, exist_ok=True)
    paths = {}
    p = os.path.join(dirpath, "h_minus_bf.dat")
    with open(p, "w") as fh:
        fh.write("# wavelength [A], cross-section [cm^2]\n")
        for x, v in zip(t["Hminus_bf_x"], t["Hminus_bf_values"]):
            fh.write(f"{float(x)!r},{float(v)!r}\n")
    paths["Hminus_bf"] = p
    p = os.path.join(dirpath, "h_minus_ff.dat")
    with open(p, "w") as fh:
        fh.write("# wavelength [A] x theta = 5040/T; 1e-26 cm^4/dyn\n")
        fh.write(", " + ",  ".join(repr(float(y)) for y in t["Hminus_ff_y"]) + "\n")
        for x, row in zip(t["Hminus_ff_x"], t["Hminus_ff_values"]):
            fh.write(f"{int(x)} " + " ".join(repr(float(v)) for v in row) + "\n")
    paths["Hminus_ff"] = p
    p = os.path.join(dirpath, "h2_plus_bf.dat")
    with open(p, "w") as fh:
        fh.write("# wavelength [nm] x T [K]; 1e-18 cm^2\n")
        fh.write("(nxn)\t" + "\t".join(str(int(y)) for y in t["H2plus_bf_y"]) + "\t\n")
        for x, row in zip(t["H2plus_bf_x"], t["H2plus_bf_values"]):
            fh.write(f"{int(x)}\t" + "\t".join(repr(float(v)) for v in row) + "\t\n")
    paths["H2plus_bf"] = p
    return paths
