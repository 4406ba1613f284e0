"""CPU model of the far-field expansion used by k_far_coeffs / k_lines (stardis_b200/csrc/k2_lines.cu): the Taylor
series of the two Lorentzian poles of the region-I profile about a tile centre, with the kernel's series-length rule
and its three-term recurrence for Im(w^k), against the closed form.  Pins the error budget quoted in DESIGN.md
(a few 1e-12 relative, all terms positive) independently of the GPU."""
import numpy as np

K1 = 27                 # SD_FAR_K + 1
RHO_INV = 3.0           # SD_FAR_RHO_INV
LOG2_RHO_INV = np.float32(1.5849625)


def _terms_needed(rho2):
    lg = -0.5 * np.log2(rho2.astype(np.float32)).astype(np.float32)
    n = np.where(lg > LOG2_RHO_INV, np.minimum(K1, (np.float32(K1) * LOG2_RHO_INV / lg + np.float32(1.02)).astype(np.int64)), K1)
    return ((n + 2) // 3) * 3  # the kernel tests the length every third term


def _series(nu, nu_c, h, nu_l, dw, y, K):
    """Sum_k C_k t^k as the kernel forms it (one pair), evaluated by Horner at t = (nu - nu_c) / h."""
    g = y * dw
    Wn = -K * dw * (0.5 / np.sqrt(np.pi)) / h
    adw = dw / np.sqrt(2.0)
    coefs = np.zeros(K1)
    rho2 = 0.0
    poles = []
    for sgn in (+1.0, -1.0):
        Dp = nu_c - (nu_l + sgn * adw)
        q = 1.0 / (Dp * Dp + g * g)
        wr, wi = Dp * (-h * q), g * (-h * q)
        rho2 = max(rho2, h * h * q)
        poles.append((wr, wi))
    nt = int(min(K1, _terms_needed(np.array([rho2]))[0]))
    for wr, wi in poles:
        a, b = wr + wr, wr * wr + wi * wi
        s, sp = wi, 0.0
        for k in range(nt):
            coefs[k] += Wn * s
            s, sp = a * s - b * sp, s
    t = (nu - nu_c) / h
    poly = np.full_like(t, coefs[K1 - 1])
    for k in range(K1 - 2, -1, -1):
        poly = poly * t + coefs[k]
    return poly, nt


def _direct(nu, nu_l, dw, y, K):
    x = (nu - nu_l) / dw
    q = x * x
    yy = y * y
    c1 = yy + 0.5
    return K * y / np.sqrt(np.pi) * (q + c1) / (q * (q + (2 * yy - 1)) + c1 * c1)   # region I, Re w K


def test_series_matches_region_one_profile_at_the_far_criterion_and_beyond():
    rng = np.random.default_rng(3)
    worst = 0.0
    lengths = set()
    for _ in range(4000):
        h = 10.0 ** rng.uniform(9.5, 12.0)
        dw = 10.0 ** rng.uniform(9.0, 10.0)
        y = 10.0 ** rng.uniform(-4.0, 1.0)
        # distance of the line from the tile centre: from exactly the far criterion out to 60 half-widths
        need = max(RHO_INV * h + dw / np.sqrt(2.0), h + max(15.0000001 - y, 0.0) * dw)
        dist = need * 10.0 ** rng.uniform(0.0, 1.2)
        nu_c = 6.0e14
        nu_l = nu_c + rng.choice([-1.0, 1.0]) * dist
        nu = nu_c + h * np.linspace(-1.0, 1.0, 33)
        approx, nt = _series(nu, nu_c, h, nu_l, dw, y, 1.0)
        exact = _direct(nu, nu_l, dw, y, 1.0)
        assert (exact > 0).all()
        worst = max(worst, np.max(np.abs(approx - exact) / exact))
        lengths.add(nt)
    assert worst < 2e-11, worst          # worst case of one expansion ~6e-12
    assert min(lengths) <= 9 and max(lengths) == K1   # the rule really shortens distant expansions


def test_series_length_rule_is_monotone_and_covers_the_full_series_at_the_far_criterion():
    rho = np.array([1 / RHO_INV, 0.25, 0.2, 0.125, 1 / 16, 1 / 32, 1 / 64, 1e-3])
    n = _terms_needed(rho * rho)
    assert n[0] == K1 and (np.diff(n) <= 0).all() and n[-1] >= 3
    # (n + 1) rho^n stays below the bound of the full series at the far criterion
    assert ((n + 1) * rho ** n <= (K1 + 1) * RHO_INV ** -float(K1) * 1.0000001).all()
