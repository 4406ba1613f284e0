"""Columnar, nu-sorted line table: the input format of the device line kernels.

The reference rebuilds its line table on every call with pandas (reset_index + three merges + sort + two range
filters, opacities_solvers/base.py:362-421).  ``ColumnarLines`` performs that selection ONCE
(``from_plasma``) into struct-of-arrays storage sorted by frequency; selecting the lines of a frequency grid is then
two binary searches and zero copies (contiguous slices).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import pandas as pd

_COLS = ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_lower", "level_energy_upper", "A_ul")


@dataclass
class LineStrength:
    """O(L) inputs from which the device fills ``alpha_line`` (L, D) itself instead of receiving the
    (L, D) table over PCIe.  ``kind`` = "vald": AlphaLineVald / AlphaLineShortlistVald (plasma/base.py:178-455),
    per-line ``ion_row, gf[, g_lo]`` + table ``n_over_u`` (n_ions, D), lower level energy = the line table's column;
    "levels": AlphaLine (plasma/base.py:130-175), per-line ``lower, upper, f_lu[, metastable_upper]`` + tables
    ``level_number_density`` (n_levels, D), ``g`` (n_levels,)."""
    kind: str
    per_line: dict
    tables: dict

    def take(self, key):
        return LineStrength(self.kind, {k: (None if v is None else v[key]) for k, v in self.per_line.items()}, self.tables)

    def nbytes(self):
        return int(sum(v.nbytes for v in self.per_line.values() if v is not None) + sum(np.asarray(v).nbytes for v in self.tables.values()))

    def run(self, ctx):
        """Launch the producer kernel on a context whose atmosphere and line table (nu, level_energy_lower) are set."""
        pl, tb = self.per_line, self.tables
        if self.kind == "vald":
            ctx.calc_alpha_line_vald(tb["n_over_u"], pl["ion_row"], pl["gf"], None, g_lo=pl.get("g_lo"))
        elif self.kind == "levels":
            ctx.calc_alpha_line_levels(tb["level_number_density"], tb["g"], pl["lower"], pl["upper"], pl["f_lu"],
                                       metastable_upper=pl.get("metastable_upper"))
        else:
            raise ValueError(f"unknown line-strength kind {self.kind!r}")

    def host_alpha(self, T, line_nu, e_low_erg):
        """The same (L, D) array in numpy, in the reference's order of operations (checker / CPU legs only)."""
        from ..constants import ALPHA_COEFFICIENT, H_CGS, KB_CGS

        pl, tb = self.per_line, self.tables
        T = np.asarray(T, dtype=np.float64)
        if self.kind == "vald":
            boltz = np.exp(np.outer(-e_low_erg, 1.0 / (T * KB_CGS)))
            n_lower = boltz * tb["n_over_u"][pl["ion_row"]]
            if pl.get("g_lo") is not None:
                n_lower = n_lower * pl["g_lo"][:, None]
            emis = 1.0 - np.exp((-H_CGS / KB_CGS) * np.outer(line_nu, 1.0 / T))
            return np.ascontiguousarray(ALPHA_COEFFICIENT * n_lower * pl["gf"][:, None] * emis)
        n_lo, n_up = tb["level_number_density"][pl["lower"]], tb["level_number_density"][pl["upper"]]
        with np.errstate(divide="ignore", invalid="ignore"):
            sef = 1.0 - ((tb["g"][pl["lower"]][:, None] * n_up) / (tb["g"][pl["upper"]][:, None] * n_lo))
        sef[n_lo == 0.0] = 0.0
        sef[np.isneginf(sef)] = 0.0
        if pl.get("metastable_upper") is not None:
            sef[(pl["metastable_upper"] != 0)[:, None] & (sef < 0)] = 0.0
        return np.ascontiguousarray(ALPHA_COEFFICIENT * n_lo * sef * pl["f_lu"][:, None])


def _strength_from_plasma(stellar_plasma, use_vald, order, n_rows):
    """O(L) producer inputs of a tardis-style plasma, permuted into the nu-sorted row order (``order`` = original row of
    every sorted row), or None when the plasma does not expose them (the (L, D) table is uploaded then).

    non-VALD (AlphaLine, plasma/base.py:143-158): ``level_number_density``, ``lines_lower_level_index``,
    ``lines_upper_level_index``, ``g``, ``metastability``, ``f_lu``.  VALD (AlphaLineVald / AlphaLineShortlistVald,
    :203-321 / :346-455): ``atomic_data.linelist_atoms``, ``ion_number_density``, ``partition_function``,
    ``ionization_data``; the electron temperatures are the model temperatures (stardis links t_electrons = t_rad)."""
    try:
        if not use_vald:
            lnd = stellar_plasma.level_number_density
            lower = np.asarray(stellar_plasma.lines_lower_level_index, dtype=np.int64)
            upper = np.asarray(stellar_plasma.lines_upper_level_index, dtype=np.int64)
            f_lu = np.asarray(getattr(stellar_plasma.f_lu, "values", stellar_plasma.f_lu), dtype=np.float64)
            if not (len(lower) == len(upper) == len(f_lu) == n_rows):
                return None
            g = np.asarray(getattr(stellar_plasma.g, "values", stellar_plasma.g), dtype=np.float64)
            meta = np.asarray(getattr(stellar_plasma.metastability, "values", stellar_plasma.metastability)).astype(np.int64)
            per_line = dict(lower=np.ascontiguousarray(lower[order]), upper=np.ascontiguousarray(upper[order]),
                            f_lu=np.ascontiguousarray(f_lu[order]), metastable_upper=np.ascontiguousarray(meta[upper][order]))
            return LineStrength("levels", per_line, dict(level_number_density=np.ascontiguousarray(lnd.values, dtype=np.float64), g=g))
        from .alpha_line_vald import prepare_vald_linelist

        adata = stellar_plasma.atomic_data
        ind, pf = stellar_plasma.ion_number_density, stellar_plasma.partition_function
        linelist = adata.linelist_atoms
        shortlist = "e_up" not in linelist.columns
        ions = np.array([(int(a), int(b)) for a, b in ind.index])
        ion_data = stellar_plasma.ionization_data
        vl = prepare_vald_linelist(linelist, ions, np.array([(int(a), int(b)) for a, b in ion_data.index]),
                                   np.asarray(ion_data.values, dtype=np.float64), int(np.max(adata.selected_atomic_numbers)),
                                   shortlist=shortlist)
        if len(vl) != n_rows:
            return None
        per_line = dict(ion_row=np.ascontiguousarray(vl.ion_row[order]), gf=np.ascontiguousarray(vl.gf[order]),
                        g_lo=None if vl.g_lo is None else np.ascontiguousarray(vl.g_lo[order]))
        return LineStrength("vald", per_line, dict(n_over_u=np.ascontiguousarray(np.asarray(ind.values, dtype=np.float64) /
                                                                                 np.asarray(pf.loc[ind.index].values, dtype=np.float64))))
    except (AttributeError, KeyError, TypeError, ValueError):
        return None


@dataclass
class ColumnarLines:
    nu: np.ndarray                    # (L,) ascending
    atomic_number: np.ndarray         # (L,) int64
    ion_number: np.ndarray            # (L,) int64, 0 = neutral
    ionization_energy: np.ndarray     # (L,) erg
    level_energy_lower: np.ndarray    # (L,) erg
    level_energy_upper: np.ndarray    # (L,) erg
    A_ul: np.ndarray                  # (L,)
    alpha_line: np.ndarray            # (L, D) cm^-1 Hz
    stark: np.ndarray | None = None   # VALD parameters
    waals: np.ndarray | None = None
    mass: np.ndarray | None = None    # (L,) g; filled from the model composition (or molecular masses)
    strength: "LineStrength | None" = None  # O(L) producer inputs: the device fills alpha_line itself when present
    _no_autoion: "ColumnarLines | None" = field(default=None, repr=False)
    _mass_src: object = field(default=None, repr=False)  # the nuclide_masses object `mass` was derived from ("given": supplied)

    def __post_init__(self):
        if self.mass is not None and self._mass_src is None:
            self._mass_src = "given"

    def __len__(self):
        return int(self.nu.shape[0])

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_arrays(cls, **cols):
        order = np.argsort(cols["nu"], kind="stable")
        out = {}
        for k, v in cols.items():
            if v is None:
                out[k] = None
                continue
            v = np.asarray(v)
            out[k] = np.ascontiguousarray(v[order])
        for k in ("atomic_number", "ion_number"):
            out[k] = out[k].astype(np.int64)
        return cls(**out)

    @classmethod
    def from_plasma(cls, stellar_plasma, use_vald=False):
        """The reference's table assembly (opacities_solvers/base.py:362-407), done once.

        non-VALD: ``plasma.lines`` joined with ``plasma.ionization_data`` (ion_number - 1) and twice with
        ``plasma.atomic_data.levels.energy``; VALD: ``plasma.lines_from_linelist`` as is.  Rows are then sorted by
        nu, as are the rows of ``plasma.alpha_line`` / ``alpha_line_from_linelist`` (the reference pairs the two
        tables positionally after sorting both by nu, :392-407)."""
        if use_vald:
            lines = stellar_plasma.lines_from_linelist
            alphas_and_nu = stellar_plasma.alpha_line_from_linelist
        else:
            lines = stellar_plasma.lines.reset_index()
            ionization_data = stellar_plasma.ionization_data.reset_index()
            ionization_data["ion_number"] -= 1
            lines = pd.merge(lines, ionization_data, how="left", on=["atomic_number", "ion_number"])
            levels_energy = stellar_plasma.atomic_data.levels.energy
            for side in ("lower", "upper"):
                lines = pd.merge(lines, levels_energy, how="left",
                                 left_on=["atomic_number", "ion_number", f"level_number_{side}"],
                                 right_on=["atomic_number", "ion_number", "level_number"],
                                 ).rename(columns={"energy": f"level_energy_{side}"})
            alphas_and_nu = stellar_plasma.alpha_line
        lines_sorted = lines.sort_values("nu")
        alphas_sorted = alphas_and_nu.sort_values("nu")
        cols = {k: pd.to_numeric(lines_sorted[k]).to_numpy() for k in _COLS}
        cols["alpha_line"] = alphas_sorted.drop(labels="nu", axis=1).to_numpy(dtype=np.float64)
        if use_vald and "stark" in lines_sorted:
            cols["stark"] = lines_sorted["stark"].to_numpy(dtype=np.float64)
            cols["waals"] = lines_sorted["waals"].to_numpy(dtype=np.float64)
        out = {k: np.ascontiguousarray(v) for k, v in cols.items()}
        out["atomic_number"] = out["atomic_number"].astype(np.int64)
        out["ion_number"] = out["ion_number"].astype(np.int64)
        out["strength"] = _strength_from_plasma(stellar_plasma, use_vald, lines_sorted.index.to_numpy(), len(lines))
        return cls(**out)

    # ------------------------------------------------------------------ selection
    def without_autoionizing(self):
        """Drop lines whose upper level lies above the ionisation energy (base.py:413-421); cached."""
        if self._no_autoion is None:
            keep = ~(self.level_energy_upper > self.ionization_energy)
            if keep.all():
                self._no_autoion = self
            else:
                self._no_autoion = ColumnarLines(**{k: (None if getattr(self, k) is None else np.ascontiguousarray(getattr(self, k)[keep]))
                                                   for k in self._data_fields()},
                                                 strength=None if self.strength is None else self.strength.take(keep))
        return self._no_autoion

    def in_range(self, nu_min, nu_max):
        """Lines with nu_min <= nu <= nu_max (pandas ``between`` is inclusive, base.py:393-396): contiguous views."""
        a = int(np.searchsorted(self.nu, nu_min, side="left"))
        b = int(np.searchsorted(self.nu, nu_max, side="right"))
        return ColumnarLines(**{k: (None if getattr(self, k) is None else getattr(self, k)[a:b]) for k in self._data_fields()},
                             strength=None if self.strength is None else self.strength.take(slice(a, b)))

    def with_masses(self, nuclide_masses):
        """mass[l] = composition.nuclide_masses.loc[atomic_number] (broadening.py:723-730)."""
        if self.mass is not None and (self._mass_src is nuclide_masses or isinstance(self._mass_src, str)):
            return self  # derived from this very composition (or supplied with the table): nothing to do
        Z = self.atomic_number
        if isinstance(nuclide_masses, pd.Series):
            table = np.full(int(max(nuclide_masses.index.max(), Z.max() if len(Z) else 1)) + 1, np.nan)
            table[nuclide_masses.index.values.astype(int)] = nuclide_masses.values
        else:
            table = np.asarray(nuclide_masses, dtype=np.float64)
        self.mass = np.ascontiguousarray(table[Z])
        self._mass_src = nuclide_masses  # another stellar model's composition recomputes the masses (Doppler widths)
        return self

    @staticmethod
    def _data_fields():
        return ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_lower", "level_energy_upper",
                "A_ul", "alpha_line", "stark", "waals", "mass")
