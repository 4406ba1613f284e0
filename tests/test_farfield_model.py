"""CPU model of the far-field expansions used by k_far_coeffs / k_s2m / k_m2l / k_lines (stardis_b200/csrc/k2_lines.cu):
the Taylor series of the two Lorentzian poles of the region-I profile about a tile centre (direct expansion), and the
multipole moments of a source tile translated to Taylor coefficients of a target tile (1-D fast multipole step), with
the kernels' series-length rule and three-term recurrences, against the closed form.  Pins the error budget quoted in
DESIGN.md (a few 1e-12 relative per pair in the worst geometry, all terms positive) independently of the GPU."""
import numpy as np

K1 = 32                 # SD_FAR_K + 1
RHO = 0.40              # SD_FAR_RHO
RHO_INV = 1.0 / RHO
LOG2_RHO_INV = np.float32(1.3219281)


def _terms_needed(rho2):
    lg = -0.5 * np.log2(rho2.astype(np.float32)).astype(np.float32)
    n = np.where(lg > LOG2_RHO_INV, np.minimum(K1, (np.float32(K1) * LOG2_RHO_INV / lg + np.float32(1.02)).astype(np.int64)), K1)
    return ((n + 3) // 4) * 4  # the kernel tests the length every fourth term


def _series(nu, nu_c, h, nu_l, dw, y, K):
    """Sum_k C_k t^k as the kernel forms it (one pair), evaluated by Horner at t = (nu - nu_c) / h."""
    g = y * dw
    An = K * dw * (0.5 / np.sqrt(np.pi))
    adw = dw / np.sqrt(2.0)
    coefs = np.zeros(K1)
    rho2 = 0.0
    poles = []
    for sgn in (+1.0, -1.0):
        Dp = nu_c - (nu_l + sgn * adw)
        q = 1.0 / (Dp * Dp + g * g)
        wr, wi = Dp * (-h * q), g * (-h * q)
        rho2 = max(rho2, h * h * q)
        poles.append((wr, wi, g * q))
    nt = int(min(K1, _terms_needed(np.array([rho2]))[0]))
    for wr, wi, vi in poles:
        a, b = wr + wr, wr * wr + wi * wi
        s, sp = vi, 0.0          # Im(v w^k): s_0 = Im v, s_-1 = Im(-1/h) = 0
        for k in range(nt):
            coefs[k] += An * s
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
        # distance of the line from the tile centre: from exactly the convergence limit out to 60 half-widths
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
    assert min(lengths) <= 12 and max(lengths) == K1   # the rule really shortens distant expansions


def test_series_length_rule_is_monotone_and_covers_the_full_series_at_the_far_criterion():
    rho = np.array([RHO, 1 / 3, 0.25, 0.2, 0.125, 1 / 16, 1 / 32, 1 / 64, 1e-3])
    n = _terms_needed(rho * rho)
    assert n[0] == K1 and (np.diff(n) <= 0).all() and n[-1] >= 3
    # (n + 1) rho^n stays below the bound of the full series at the far criterion
    assert ((n + 1) * rho ** n <= (K1 + 1) * RHO_INV ** -float(K1) * 1.0000001).all()


# ---- multipole moments of a source tile + tile-to-tile translation (k_s2m, k_m2l) ---------------------------------
def _moments(nu_l, dw, y, K, c_s, scale):
    """M_k = A Im(u+^k + u-^k), k = 1..K1, u = (pole - c_s) / scale, by the kernel's three-term recurrence."""
    An = K * dw * (0.5 / np.sqrt(np.pi))
    g, adw = y * dw, dw / np.sqrt(2.0)
    m = np.zeros(K1)
    for sgn in (+1.0, -1.0):
        ur, ui = (nu_l - c_s + sgn * adw) / scale, g / scale
        a, b = ur + ur, ur * ur + ui * ui
        s, sp = ui, 0.0
        for k in range(K1):
            m[k] += An * s
            s, sp = a * s - b * sp, s
    return m


def _translate(m, c_s, scale, c_t, h_t, kmax=K1):
    """L_n = (b^n / d) sum_k C(n + k, n) a^k M_k with the row recurrence of k_m2l."""
    d = c_t - c_s
    a, b = scale / d, -h_t / d
    L = np.zeros(K1)
    for n in range(K1):
        coef = b ** n / d * (n + 1) * a
        acc = 0.0
        for k in range(kmax):
            acc += coef * m[k]
            coef *= a * (n + k + 2) / (k + 2)
        L[n] = acc
    return L


def test_multipole_translation_matches_region_one_profile():
    rng = np.random.default_rng(11)
    worst, used, shortest = 0.0, 0, K1
    for _ in range(3000):
        h_s = 10.0 ** rng.uniform(9.5, 12.0)
        r1, r2 = rng.uniform(0.8, 1.25, 2)           # neighbouring tiles of a smooth, non-uniform grid
        h_mid, h_t = h_s * r1, h_s * r1 * r2
        gap = rng.integers(0, 6)                      # further tiles in between (index distance 2 + gap)
        c_s = 6.0e14
        sgn = rng.choice([-1.0, 1.0])
        c_t = c_s + sgn * (h_s + 2.0 * h_mid * (1 + gap) + h_t)
        dw = rng.uniform(0.0, 0.12) * h_s * np.sqrt(2.0) + 1.0
        g = rng.uniform(0.0, max(0.2 * h_s - dw / np.sqrt(2.0), 0.0))   # overhang of the poles <= 0.2 h (SD_FAR_OVERHANG)
        y = g / dw
        nu_l = c_s + rng.uniform(-1.0, 1.0) * h_s
        q_mul = np.hypot(abs(nu_l - c_s) + dw / np.sqrt(2.0), g) / (abs(c_t - c_s) - h_t)
        q_loc = h_t / (abs(c_t - nu_l) - dw / np.sqrt(2.0))
        if max(q_mul, q_loc) > RHO:                   # level_ok() of k_build_records rejects these at this level
            continue
        rr = 1.2 * h_s / (abs(c_t - c_s) - h_t)
        kmax = int(min(K1, _terms_needed(np.array([rr * rr]))[0]))
        shortest = min(shortest, kmax)
        nu = c_t + h_t * np.linspace(-1.0, 1.0, 33)
        L = _translate(_moments(nu_l, dw, y, 1.0, c_s, h_s), c_s, h_s, c_t, h_t, kmax)
        t = (nu - c_t) / h_t
        poly = np.full_like(t, L[K1 - 1])
        for n in range(K1 - 2, -1, -1):
            poly = poly * t + L[n]
        exact = _direct(nu, nu_l, dw, y, 1.0)
        worst = max(worst, np.max(np.abs(poly - exact) / exact))
        used += 1
    assert used > 1500
    assert worst < 2e-11, worst
    assert shortest <= 16   # distant source tiles need far fewer moments


def _translate_up(m, c_s, scale_s, c_p, scale_p):
    """k_m2m: moments about (c_s, scale_s) -> moments about the parent's (c_p, scale_p):
    M'_k = sum_{j=1..k} C(k, j) r^j delta^(k-j) M_j,  r = scale_s / scale_p,  delta = (c_s - c_p) / scale_p,
    with the kernel's running coefficient C(k, j) r^j."""
    r, delta = scale_s / scale_p, (c_s - c_p) / scale_p
    dp = delta ** np.arange(K1 + 1)
    out = np.zeros(K1)
    for kk in range(1, K1 + 1):
        cf, rem, acc = kk * r, float(kk - 1), 0.0
        for j in range(1, kk + 1):
            acc += cf * dp[kk - j] * m[j - 1]
            cf *= (r * (1.0 / (j + 1))) * rem
            rem -= 1.0
        out[kk - 1] = acc
    return out


def test_moment_translation_to_the_parent_tile_is_exact():
    """Moments of a child tile translated to its parent equal the moments taken about the parent directly (a finite
    binomial sum: no truncation, only rounding), for every child position of a branching-8 hierarchy."""
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(400):
        h_p = 10.0 ** rng.uniform(10.0, 12.5)
        c_p = 6.0e14
        child = rng.integers(0, 8)
        h_s = h_p / 8.0 * rng.uniform(0.95, 1.05)                  # smooth non-uniform grid
        c_s = c_p + h_p * (-1.0 + (2 * child + 1) / 8.0)
        dw = rng.uniform(0.0, 0.1) * h_s + 1.0
        y = rng.uniform(0.0, 0.1) * h_s / dw
        nu_l = c_s + rng.uniform(-1.0, 1.0) * h_s
        direct = _moments(nu_l, dw, y, 1.0, c_p, h_p)
        up = _translate_up(_moments(nu_l, dw, y, 1.0, c_s, h_s), c_s, h_s, c_p, h_p)
        scale = np.max(np.abs(direct))
        worst = max(worst, np.max(np.abs(up - direct)) / scale)
    assert worst < 1e-13, worst
