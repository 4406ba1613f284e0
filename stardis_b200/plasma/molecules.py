"""Molecular plasma state and molecular line strengths (stardis/plasma/molecules.py:16-445) -- the inputs of the
molecular branch of the line-opacity kernels.

Molecule number densities (Barklem & Collet 2016 equilibrium constants) and partition functions are O(molecules x D)
numbers and are formed on the host; the (L_mol, D) line-strength table is filled on the DEVICE by the same kernel as the
atomic VALD lists (``sd_calc_alpha_line_vald``: the formula of ``AlphaLineValdMolecule`` / ``AlphaLineShortlistValdMolecule``
is that of their atomic twins with N_molecule / U_molecule in place of N_ion / U_ion), so nothing of size L x D crosses
PCIe.  Inputs are the tables the reference reads from ``atomic_data.molecule_data`` / ``atomic_data.linelist_molecules``.
"""
from __future__ import annotations

import re

import numpy as np
import pandas as pd

from .. import constants as const
from .alpha_line_vald import EV_ERG, ValdLines

ELEMENT_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y "
                   "Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os "
                   "Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U").split()
_ION = re.compile(r"([A-Z][a-z]?)(\+*)(\-*)")
PA_IN_CGS = 10.0  # 1 Pa = 10 dyn cm^-2


def split_ion(name):
    """'Ti' -> (22, 0), 'C+' -> (6, 1), 'H-' -> (1, -1): atomic number and charge of a constituent (molecules.py:145-159)."""
    m = _ION.match(str(name))
    if m is None or m.group(1) not in ELEMENT_SYMBOLS:
        raise ValueError(f"cannot parse the molecule constituent {name!r}")
    return ELEMENT_SYMBOLS.index(m.group(1)) + 1, len(m.group(2)) - len(m.group(3))


def molecule_number_density(ion_number_density, t_electrons, molecule_data):
    """``MoleculeIonNumberDensity`` (molecules.py:16-143) -> (molecule_number_density, molecule_ion_map) DataFrames.

    ``molecule_data.dissociation_energies``: rows = molecules with the constituents ``Ion1`` / ``Ion2`` as strings;
    ``molecule_data.equilibrium_constants``: log10 of the pressure equilibrium constant [Pa] on a temperature grid
    (columns).  Per molecule: cubic spline in T (extrapolating), K_p -> K_n = K_p / (k_B T), then the closed-form root of
    the dissociation equilibrium -- n = (K + 4 n1 - sqrt(K (K + 8 n1))) / 8 for a homonuclear pair, otherwise
    n = (K + n1 + n2 - sqrt(K^2 + 2 K (n1 + n2) + (n1 - n2)^2)) / 2 -- clipped at 0.  Molecules with a negative ion or
    with a constituent element that is not in the plasma get zero density (the reference logs a warning)."""
    from scipy.interpolate import CubicSpline

    try:
        diss = molecule_data.dissociation_energies
    except AttributeError:
        raise ValueError("No molecular dissociation energies found in atomic data. Use Carsus to generate atomic data with "
                         "the Barklem and Collet 2016 data.") from None
    eq = molecule_data.equilibrium_constants
    T = np.asarray(t_electrons, dtype=np.float64)
    grid = eq.columns.values.astype(np.float64)
    present = set(ion_number_density.index.get_level_values(0).unique())
    dens = np.zeros((len(eq), T.size))
    parts = {mol: (split_ion(row.Ion1), split_ion(row.Ion2)) for mol, row in diss.iterrows()}
    for mol, ((z1, q1), (z2, q2)) in parts.items():
        if q1 == -1 or q2 == -1 or z1 not in present or z2 not in present:
            continue
        n1 = np.asarray(ion_number_density.loc[z1, q1], dtype=np.float64)
        n2 = np.asarray(ion_number_density.loc[z2, q2], dtype=np.float64)
        log_kp = CubicSpline(grid, eq.loc[mol].values, extrapolate=True)(T)
        K = (10.0 ** log_kp) * PA_IN_CGS / (const.KB_CGS * T)
        if (z1, q1) == (z2, q2):
            n = (1 / 8) * ((-((K * (K + 8 * n1)) ** 0.5)) + K + 4 * n1)
        else:
            n = 0.5 * (-np.sqrt(K ** 2 + 2 * K * (n1 + n2) + (n1 - n2) ** 2) + K + n1 + n2)
        dens[eq.index.get_loc(mol)] = np.maximum(n, 0)
    density_df = pd.DataFrame(dens, index=eq.index, columns=ion_number_density.columns)
    ion_map = pd.DataFrame({"Ion1": [parts[m][0][0] for m in diss.index], "Ion2": [parts[m][1][0] for m in diss.index]}, index=diss.index)
    return density_df, ion_map


def molecule_partition_function(t_electrons, molecule_data):
    """``MoleculePartitionFunction`` (molecules.py:162-191): linear interpolation of the tabulated partition functions."""
    pf = molecule_data.partition_functions
    grid = pf.columns.values.astype(np.float64)
    T = np.asarray(t_electrons, dtype=np.float64)
    return pd.DataFrame(np.array([np.interp(T, grid, pf.loc[m].values) for m in pf.index]), index=pf.index)


def prepare_molecule_linelist(linelist_molecules, molecule_index, shortlist=False):
    """O(L) host preparation of ``AlphaLineValdMolecule.calculate`` (molecules.py:234-262, 294-318) and of the short-list
    variant (:372-389, 440-445): line frequency, gf (/ g_lo for the long format), lower level energy and, per line, the row
    of its molecule in the (molecule, depth) tables.  Rows keep the linelist's own order, as in the reference."""
    col = lambda k: np.asarray(linelist_molecules[k].values if hasattr(linelist_molecules[k], "values") else linelist_molecules[k])
    lam = col("wavelength").astype(np.float64)
    log_gf = col("log_gf").astype(np.float64)
    e_low = col("e_low").astype(np.float64)
    nu = const.C_CGS / (lam * 1e-8)
    if shortlist:
        e_up = (e_low * EV_ERG + (const.H_CGS * const.C_CGS) / (lam * 1e-8)) / EV_ERG
        g_lo, gf = None, 10 ** log_gf
    else:
        e_up = col("e_up").astype(np.float64)
        g_lo = col("j_lo").astype(np.float64) * 2 + 1
        gf = 10 ** log_gf / g_lo
    row_of = {str(m): i for i, m in enumerate(molecule_index)}
    try:
        rows = np.array([row_of[str(m)] for m in col("molecule")], dtype=np.int64)
    except KeyError as e:
        raise ValueError(f"linelist molecule {e.args[0]!r} has no row in the molecule number density table") from None
    return ValdLines(molecule=col("molecule").astype(str), nu=nu, gf=gf, g_lo=g_lo, e_low_erg=e_low * EV_ERG, ion_row=rows,
                     level_energy_lower=e_low * EV_ERG, level_energy_upper=e_up * EV_ERG, A_ul=10 ** col("rad").astype(np.float64))


def alpha_line_vald_molecule(ctx, lines, molecule_number_density_, molecule_partition_function_):
    """Fills ``molecule_alpha_line_from_linelist`` (L_mol, D) in HBM from the prepared O(L) columns (the atomic producer
    kernel with N_molecule / U_molecule); ``ctx`` needs ``set_atmosphere`` (electron temperatures = the atmosphere's).
    ``ctx.get(BUF_LINE_STRENGTH)`` copies the table back; K1/K2 of the molecular branch read it in place."""
    n = np.asarray(getattr(molecule_number_density_, "values", molecule_number_density_), dtype=np.float64)
    u_ = np.asarray(getattr(molecule_partition_function_, "values", molecule_partition_function_), dtype=np.float64)
    ctx.set_lines(lines.nu, None, mass=np.ones(len(lines)))
    ctx.calc_alpha_line_vald(n / u_, lines.ion_row, lines.gf, lines.e_low_erg, g_lo=lines.g_lo)
    return ctx
