"""Test helpers: recreate the reference's data files (cross-section tables, a MARCS .mod file) from the numeric
fixtures, so that nothing under /root/reference is needed at test time."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def write_table_files(dirpath):
    """Write the three cross-section tables in the text formats the reference ships (stardis/data/*.dat)."""
    from stardis_b200.synthetic import write_cross_section_files

    return write_cross_section_files(dirpath)


def write_marcs_mod(path, name="sun"):
    """A plane-parallel MARCS .mod file with the structure columns of the fixture atmosphere."""
    z = np.load(os.path.join(os.path.dirname(GOLDEN), "..", "benchdata", "atmospheres.npz"))
    depth, T, pe, pg, rho = (z[f"{name}_{k}"][::-1] for k in ("depth", "t", "pe", "pg", "density"))  # surface first
    logA = z[f"{name}_logA"]
    n = len(T)
    with open(path, "w") as fh:
        fh.write(f"{name}\n")
        fh.write(f"  {float(z[f'{name}_teff']):.0f}.      Teff [K].         Last iteration; yyyymmdd=20080519\n")
        fh.write("  6.3157E+10 Flux [erg/cm2/s]\n  2.7542E+04 Surface gravity [cm/s2]\n")
        fh.write(f"  {float(z[f'{name}_vmic_kms']):.1f}        Microturbulence parameter [km/s]\n")
        fh.write("  0.0        No mass for plane-parallel models\n +0.00 +0.00 Metallicity [Fe/H] and [alpha/Fe]\n")
        fh.write("  1.0000E+00 1 cm radius for plane-parallel models\n  2.0783E-22 Luminosity [Lsun] FOR A RADIUS OF 1 cm!\n")
        fh.write("  1.50 8.00 0.076 0.00 are the convection parameters: alpha, nu, y and beta\n")
        fh.write("  0.73826 0.24954 1.22E-02 are X, Y and Z, 12C/13C=89 (=solar)\n")
        fh.write("Logarithmic chemical number abundances, H always 12.00\n")
        for i in range(0, len(logA), 10):
            fh.write(" ".join(f"{a:7.2f}" for a in logA[i:i + 10]) + "\n")
        fh.write(f"  {n} Number of depth points\nModel structure\n")
        fh.write(" k lgTauR  lgTau5    Depth     T        Pe          Pg         Prad       Pturb\n")
        for k in range(n):
            fh.write(f"{k + 1:3d} {-5 + 0.1 * k:5.2f} {-4.9 + 0.1 * k:7.4f} {depth[k]: .3E} {T[k]:7.1f} {pe[k]: .4E} {pg[k]: .4E} "
                     f" 1.4884E+00  0.0000E+00\n")
        fh.write(" k lgTauR    KappaRoss   Density   Mu      Vconv   Fconv/F      RHOX\n")
        for k in range(n):
            fh.write(f"{k + 1:3d} {-5 + 0.1 * k:5.2f}  1.0000E-03 {rho[k]: .4E} 1.257  0.000E+00 0.00000  1.000000E-02\n")
        fh.write("Assorted logarithmic partial pressures\n")
    return path
