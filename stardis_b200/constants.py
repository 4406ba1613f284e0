"""CODATA-2018 CGS constants (the astropy 6.1 values the reference's environment pins) and the derived
constants of the reference's hot-path modules."""
import numpy as np

C_CGS = 2.99792458e10
H_CGS = 6.62607015e-27
KB_CGS = 1.380649e-16
E_ESU = 4.803204712570263e-10
ME_CGS = 9.1093837015e-28
MP_CGS = 1.67262192369e-24
AMU_CGS = 1.66053906660e-24
A0_CGS = 5.29177210903e-9
RYD_CGS = 109737.31568160
SIGMA_T_CGS = 6.6524587321e-25
EV_ERG = 1.602176634e-12

# opacities_solvers/base.py:20-34
BF_CONSTANT = 64 * np.pi**4 * E_ESU**10 * ME_CGS / (3 * np.sqrt(3) * C_CGS * H_CGS**6)
FF_CONSTANT = 4 / (3 * H_CGS * C_CGS) * E_ESU**6 * np.sqrt(2 * np.pi / (3 * ME_CGS**3 * KB_CGS))
RYDBERG_FREQUENCY = C_CGS * RYD_CGS
# broadening.py:20
RYDBERG_ENERGY = H_CGS * C_CGS * RYD_CGS
ALPHA_COEFFICIENT = (np.pi * E_ESU**2) / (ME_CGS * C_CGS)  # stardis/plasma/base.py:35
