from .base import (calc_alphas, calc_alpha_file, calc_alpha_bf, calc_alpha_ff, calc_alpha_rayleigh,  # noqa: F401
                   calc_alpha_electron, calc_alpha_line_at_nu, calc_molecular_alpha_line_at_nu, calc_alan_entries)
from .broadening import calculate_broadening, calculate_molecule_broadening  # noqa: F401
from .voigt import faddeeva, voigt_profile  # noqa: F401
