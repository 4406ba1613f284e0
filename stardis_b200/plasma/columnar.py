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
    _no_autoion: "ColumnarLines | None" = field(default=None, repr=False)

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
                                                   for k in self._data_fields()})
        return self._no_autoion

    def in_range(self, nu_min, nu_max):
        """Lines with nu_min <= nu <= nu_max (pandas ``between`` is inclusive, base.py:393-396): contiguous views."""
        a = int(np.searchsorted(self.nu, nu_min, side="left"))
        b = int(np.searchsorted(self.nu, nu_max, side="right"))
        return ColumnarLines(**{k: (None if getattr(self, k) is None else getattr(self, k)[a:b]) for k in self._data_fields()})

    def with_masses(self, nuclide_masses):
        """mass[l] = composition.nuclide_masses.loc[atomic_number] (broadening.py:723-730)."""
        if self.mass is not None:
            return self
        Z = self.atomic_number
        if isinstance(nuclide_masses, pd.Series):
            table = np.full(int(max(nuclide_masses.index.max(), Z.max() if len(Z) else 1)) + 1, np.nan)
            table[nuclide_masses.index.values.astype(int)] = nuclide_masses.values
        else:
            table = np.asarray(nuclide_masses, dtype=np.float64)
        self.mass = np.ascontiguousarray(table[Z])
        return self

    @staticmethod
    def _data_fields():
        return ("nu", "atomic_number", "ion_number", "ionization_energy", "level_energy_lower", "level_energy_upper",
                "A_ul", "alpha_line", "stark", "waals", "mass")
