"""VALD line strengths on the device -- the producer of ``alpha_line`` immediately upstream of the hot path
(SURVEY 8f rank 1): ``AlphaLineVald`` (stardis/plasma/base.py:178-321) and ``AlphaLineShortlistVald`` (:324-455).

The reference builds ``alpha_line_from_linelist`` (L, D) with pandas / astropy on the host; the B200 path then has to
upload it (L·D·8 bytes, the largest host->device transfer of a run).  Here the host only prepares O(L) columns (line
frequency, gf, statistical weight, lower level energy, row of the (ion, depth) table) and ``k_alpha_line_vald`` fills
the (L, D) array in HBM, where K1/K2 read it.  Same truncation / auto-ionisation filtering and the same line columns
as the reference (``lines_from_linelist``)."""
from __future__ import annotations

import numpy as np

from .. import constants as const

EV_ERG = 1.602176634e-12


class ValdLines:
    """Prepared VALD linelist: nu-sorted line columns for K1/K2 plus the O(L) inputs of ``k_alpha_line_vald``."""

    def __init__(self, **cols):
        self.__dict__.update(cols)

    def __len__(self):
        return int(self.nu.shape[0])


def _col(linelist, name):
    v = linelist[name]
    return np.asarray(getattr(v, "values", v))


def prepare_vald_linelist(linelist, ions, ionization_index, ionization_energy, max_atomic_number, shortlist=False):
    """Host preparation = everything of ``calculate`` that is O(L) (plasma/base.py:216-240, 264-270, 297-318).

    ``linelist``: mapping / DataFrame with ``atomic_number, ion_charge, wavelength [A], log_gf, e_low [eV], rad, stark,
    waals`` (+ ``e_up [eV], j_lo`` for the long format).  ``ions``: (n_ions, 2) (atomic_number, charge) rows of the
    (ion, depth) tables passed to ``alpha_line_vald``.  ``ionization_index``: (n, 2) (atomic_number, ion_number) with the
    tardis convention ion_number = charge + 1; ``ionization_energy`` [erg].  Returns ``ValdLines`` in the reference's row
    order (``alpha_line_from_linelist`` / ``lines_from_linelist`` order: the linelist's own)."""
    Z_all = _col(linelist, "atomic_number").astype(np.int64)
    keep = Z_all <= int(max_atomic_number)                                      # :235-237 / :375-377
    Z = Z_all[keep]
    q = _col(linelist, "ion_charge").astype(np.int64)[keep]
    lam = _col(linelist, "wavelength").astype(np.float64)[keep]
    log_gf = _col(linelist, "log_gf").astype(np.float64)[keep]
    e_low = _col(linelist, "e_low").astype(np.float64)[keep]
    nu = const.C_CGS / (lam * 1e-8)                                             # :268-270
    if shortlist:
        e_up = (e_low * EV_ERG + (const.H_CGS * const.C_CGS) / (lam * 1e-8)) / EV_ERG  # :380-387
        g_lo, gf = None, 10 ** log_gf
    else:
        e_up = _col(linelist, "e_up").astype(np.float64)[keep]
        g_lo = _col(linelist, "j_lo").astype(np.float64)[keep] * 2 + 1         # :240
        gf = 10 ** log_gf / g_lo                                                # :264-266
    row_of = {(int(a), int(b)): i for i, (a, b) in enumerate(np.asarray(ions))}
    try:
        ion_row = np.array([row_of[(int(a), int(b))] for a, b in zip(Z, q)], dtype=np.int64)
    except KeyError as e:
        raise ValueError(f"linelist ion {e.args[0]} has no row in the ion number density table") from None
    ion_e = {(int(a), int(b) - 1): float(e) for (a, b), e in zip(np.asarray(ionization_index), np.asarray(ionization_energy))}
    eion = np.array([ion_e.get((int(a), int(b)), np.nan) for a, b in zip(Z, q)])   # left merge :299-305
    cols = dict(atomic_number=Z, ion_number=q, nu=nu, gf=gf, g_lo=g_lo, e_low_erg=e_low * EV_ERG, ion_row=ion_row,
                level_energy_lower=e_low * EV_ERG, level_energy_upper=e_up * EV_ERG,
                A_ul=10 ** _col(linelist, "rad").astype(np.float64)[keep], ionization_energy=eion,
                stark=_col(linelist, "stark").astype(np.float64)[keep], waals=_col(linelist, "waals").astype(np.float64)[keep])
    if not shortlist:  # auto-ionising lines cannot be broadened (:316-318)
        valid = cols["level_energy_upper"] < cols["ionization_energy"]
        cols = {k: (v[valid] if v is not None else None) for k, v in cols.items()}
    return ValdLines(**cols)


def alpha_line_vald(ctx, lines, ion_number_density, partition_function, masses=None):
    """Uploads the prepared O(L) columns, runs ``k_alpha_line_vald`` and leaves ``alpha_line`` (L, D) in HBM as the line
    table's strengths (``ctx.get(BUF_LINE_STRENGTH)`` copies it back).  ``ctx`` needs ``set_atmosphere`` first; the
    electron temperatures are the atmosphere's.  ``ion_number_density`` / ``partition_function``: (n_ions, D) arrays (or
    DataFrames) in the row order given to ``prepare_vald_linelist``."""
    n = np.asarray(getattr(ion_number_density, "values", ion_number_density), dtype=np.float64)
    u_ = np.asarray(getattr(partition_function, "values", partition_function), dtype=np.float64)
    mass = masses if masses is not None else np.ones(len(lines))
    ctx.set_lines(lines.nu, None, mass=mass, atomic_number=lines.atomic_number, ion_number=lines.ion_number,
                  ionization_energy=lines.ionization_energy, level_energy_upper=lines.level_energy_upper,
                  level_energy_lower=lines.level_energy_lower, A_ul=lines.A_ul, stark=lines.stark, waals=lines.waals)
    ctx.calc_alpha_line_vald(n / u_, lines.ion_row, lines.gf, lines.e_low_erg, g_lo=lines.g_lo)   # N / U as :250-254
    return ctx
