// sd_math.cuh -- scalar fp64 building blocks shared by all kernels (device) and by the host-side
// arithmetic harness used in the CPU tests (tests/host_math_harness.cu compiles this header for the host).
//
// Everything here is real arithmetic: the reference evaluates the Humlicek W4 Faddeeva approximation in
// complex128 (stardis/radiation_field/opacities/opacities_solvers/voigt.py:17-86) and then keeps only the
// real part (voigt.py:149); we expand the complex products/quotients by hand and never form Im(w).
// The region tests use the SAME comparisons, literals and operand order as voigt.py:37-44 so that the
// classification (and the NaN behaviour: NaN -> region IV) is identical.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SD_HD __host__ __device__ __forceinline__
#else
#define SD_HD static inline
#endif

namespace sdm {

// CODATA-2018 CGS constants = astropy 6.1 (pinned by the reference's lock files)
constexpr double C_CGS = 2.99792458e10;
constexpr double H_CGS = 6.62607015e-27;
constexpr double KB_CGS = 1.380649e-16;
constexpr double E_ESU = 4.803204712570263e-10;
constexpr double A0_CGS = 5.29177210903e-9;
constexpr double MP_CGS = 1.67262192369e-24;
constexpr double AMU_CGS = 1.66053906660e-24;
constexpr double RYD_CGS = 109737.31568160;
constexpr double PI = 3.141592653589793;
constexpr double SQRT_PI = 1.7724538509055159;
constexpr double INV_SQRT_PI = 0.5641895835477563;
constexpr double SQRT_PI_PI = 5.568327996831707;  // sqrt(pi)*pi, voigt.py:148
constexpr double RYDBERG_ENERGY = (H_CGS * C_CGS) * RYD_CGS;  // broadening.py:20

// ---------------------------------------------------------------------------- reciprocal
// 1/d to ~1 ulp for normal, finite, non-zero d: MUFU.RCP64H seed (~2^-20) + one cubically convergent
// step (r(1 + e + e^2), e = 1 - d r).  Only used where d is a sum of squares bounded away from zero.
SD_HD double rcp_fast(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    double p = fma(e, e, e);
    return fma(r, p, r);
#else
    return 1.0 / d;
#endif
}

// Same seed with ONE Newton step r(1 + e): relative error = e^2 <= 1e-12 (seed error measured on B200: 9.9e-7).
// Default in the far-wing loop: every term is positive, so alpha_line inherits at most this relative error, four
// orders of magnitude inside the 1e-8 parity tolerance.
SD_HD double rcp_fast2(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}

// ---------------------------------------------------------------------------- Humlicek W4, Re(w)
// Region I (s > 15): w = (i/sqrt(pi)) z / (z^2 - 1/2).  With q = x^2:
//   Re w = y (q + y^2 + 1/2) / (sqrt(pi) ((q - y^2 - 1/2)^2 + 4 q y^2))
//        = y (q + c1) / (sqrt(pi) (q (q + b) + c)),  c1 = y^2 + 1/2, b = 2 y^2 - 1, c = c1^2.
SD_HD double region1_re(double q, double y) {
    double yy = y * y;
    double c1 = yy + 0.5;
    double den = fma(q, q + (2.0 * yy - 1.0), c1 * c1);
    return INV_SQRT_PI * y * (q + c1) / den;
}

// Regions II-IV in real arithmetic.  t = y - i x (voigt.py:33), u = t^2.
SD_HD double region2_re(double x, double y) {
    // w = i z (z^2/sqrt(pi) - 1.4104739589) / (0.75 + z^2 (z^2 - 3))  ->  Re w = -Im(N/Dn)
    double qr = x * x - y * y, qi = 2.0 * x * y;
    double pr = qr / SQRT_PI - 1.4104739589, pi_ = qi / SQRT_PI;
    double nr = x * pr - y * pi_, ni = x * pi_ + y * pr;
    double er = qr - 3.0;
    double dr = 0.75 + (qr * er - qi * qi), di = qr * qi + qi * er;
    return (nr * di - ni * dr) / (dr * dr + di * di);
}

SD_HD void cmul_add(double &ar, double &ai, double tr, double ti, double c) {
    // (ar + i ai) <- (ar + i ai) (tr + i ti) + c
    double r = fma(ar, tr, fma(-ai, ti, c));
    double i = fma(ar, ti, ai * tr);
    ar = r;
    ai = i;
}

SD_HD double region3_re(double x, double y) {
    double tr = y, ti = -x;
    double nr = fma(0.5642236, tr, 3.778987), ni = 0.5642236 * ti;
    cmul_add(nr, ni, tr, ti, 11.96482);
    cmul_add(nr, ni, tr, ti, 20.20933);
    cmul_add(nr, ni, tr, ti, 16.4955);
    double dr = tr + 6.699398, di = ti;
    cmul_add(dr, di, tr, ti, 21.69274);
    cmul_add(dr, di, tr, ti, 39.27121);
    cmul_add(dr, di, tr, ti, 38.82363);
    cmul_add(dr, di, tr, ti, 16.4955);
    return (nr * dr + ni * di) / (dr * dr + di * di);
}

SD_HD void cmul_rsub(double &ar, double &ai, double ur, double ui, double c) {
    // (ar + i ai) <- c - u (ar + i ai)
    double r = c - (ur * ar - ui * ai);
    double i = -(ur * ai + ui * ar);
    ar = r;
    ai = i;
}

SD_HD double region4_re(double x, double y) {
    double tr = y, ti = -x;
    double ur = y * y - x * x, ui = -2.0 * x * y;
    double pr = 1.320522 - ur * 0.56419, pi_ = -ui * 0.56419;
    cmul_rsub(pr, pi_, ur, ui, 35.7668);
    cmul_rsub(pr, pi_, ur, ui, 219.031);
    cmul_rsub(pr, pi_, ur, ui, 1540.787);
    cmul_rsub(pr, pi_, ur, ui, 3321.99);
    cmul_rsub(pr, pi_, ur, ui, 36183.31);
    double nr = tr * pr - ti * pi_, ni = tr * pi_ + ti * pr;
    double qr = 1.84144 - ur, qi = -ui;
    cmul_rsub(qr, qi, ur, ui, 61.5704);
    cmul_rsub(qr, qi, ur, ui, 364.219);
    cmul_rsub(qr, qi, ur, ui, 2186.18);
    cmul_rsub(qr, qi, ur, ui, 9022.23);
    cmul_rsub(qr, qi, ur, ui, 24322.8);
    cmul_rsub(qr, qi, ur, ui, 32066.6);
    double quot = (nr * qr + ni * qi) / (qr * qr + qi * qi);
    return exp(ur) * cos(ui) - quot;
}

// Region index 0..3 with the reference's comparisons and order (voigt.py:37-44).
SD_HD int humlicek_region(double x, double y) {
    double ax = fabs(x);
    double s = ax + y;
    if (s > 15.0) return 0;
    if (s > 5.5) return 1;
    if (y >= 0.195 * ax - 0.176) return 2;
    return 3;
}

SD_HD double humlicek_re(double x, double y) {
    switch (humlicek_region(x, y)) {
        case 0: return region1_re(x * x, y);
        case 1: return region2_re(x, y);
        case 2: return region3_re(x, y);
        default: return region4_re(x, y);
    }
}

// Full complex value (elementwise API twin of voigt.py:89-110 only; never used in the line kernel).
SD_HD void humlicek_complex(double x, double y, double &wr, double &wi) {
    double tr = y, ti = -x;
    int reg = humlicek_region(x, y);
    if (reg == 0) {
        double ar = x * x - y * y - 0.5, ai = 2.0 * x * y;  // z^2 - 1/2
        double den = ar * ar + ai * ai;
        // i z / (z^2 - 1/2) = (-y + i x)(ar - i ai)/den
        wr = INV_SQRT_PI * (-y * ar + x * ai) / den;
        wi = INV_SQRT_PI * (x * ar + y * ai) / den;
    } else if (reg == 1) {
        double qr = x * x - y * y, qi = 2.0 * x * y;
        double pr = qr / SQRT_PI - 1.4104739589, pi_ = qi / SQRT_PI;
        double nr = x * pr - y * pi_, ni = x * pi_ + y * pr;
        double er = qr - 3.0;
        double dr = 0.75 + (qr * er - qi * qi), di = qr * qi + qi * er;
        double den = dr * dr + di * di;
        double fr = (nr * dr + ni * di) / den, fi = (ni * dr - nr * di) / den;  // N/Dn
        wr = -fi;
        wi = fr;
    } else if (reg == 2) {
        double nr = fma(0.5642236, tr, 3.778987), ni = 0.5642236 * ti;
        cmul_add(nr, ni, tr, ti, 11.96482);
        cmul_add(nr, ni, tr, ti, 20.20933);
        cmul_add(nr, ni, tr, ti, 16.4955);
        double dr = tr + 6.699398, di = ti;
        cmul_add(dr, di, tr, ti, 21.69274);
        cmul_add(dr, di, tr, ti, 39.27121);
        cmul_add(dr, di, tr, ti, 38.82363);
        cmul_add(dr, di, tr, ti, 16.4955);
        double den = dr * dr + di * di;
        wr = (nr * dr + ni * di) / den;
        wi = (ni * dr - nr * di) / den;
    } else {
        double ur = y * y - x * x, ui = -2.0 * x * y;
        double pr = 1.320522 - ur * 0.56419, pi_ = -ui * 0.56419;
        cmul_rsub(pr, pi_, ur, ui, 35.7668);
        cmul_rsub(pr, pi_, ur, ui, 219.031);
        cmul_rsub(pr, pi_, ur, ui, 1540.787);
        cmul_rsub(pr, pi_, ur, ui, 3321.99);
        cmul_rsub(pr, pi_, ur, ui, 36183.31);
        double nr = tr * pr - ti * pi_, ni = tr * pi_ + ti * pr;
        double qr = 1.84144 - ur, qi = -ui;
        cmul_rsub(qr, qi, ur, ui, 61.5704);
        cmul_rsub(qr, qi, ur, ui, 364.219);
        cmul_rsub(qr, qi, ur, ui, 2186.18);
        cmul_rsub(qr, qi, ur, ui, 9022.23);
        cmul_rsub(qr, qi, ur, ui, 24322.8);
        cmul_rsub(qr, qi, ur, ui, 32066.6);
        double den = qr * qr + qi * qi;
        double e = exp(ur);
        wr = e * cos(ui) - (nr * qr + ni * qi) / den;
        wi = e * sin(ui) - (ni * qr - nr * qi) / den;
    }
}

// voigt_profile (voigt.py:113-150): x = dnu/dw and y = (gamma/(sqrt(pi) pi))/dw are two IEEE divisions
// (complex / real in numba); phi = Re w / (sqrt(pi) dw).
SD_HD double voigt_profile(double dnu, double dw, double gamma) {
    double x = dnu / dw;
    double y = (gamma / SQRT_PI_PI) / dw;
    return humlicek_re(x, y) / (SQRT_PI * dw);
}

// ---------------------------------------------------------------------------- broadening
SD_HD double doppler_width(double nu_line, double T, double mass, double vmic) {  // broadening.py:32-66
    return nu_line / C_CGS * sqrt(2.0 * KB_CGS * T / mass + vmic * vmic);
}
SD_HD double n_effective(double z_eff, double e_ion, double e_level) {  // broadening.py:114-137
    return sqrt(RYDBERG_ENERGY / (e_ion - e_level)) * z_eff;
}
SD_HD double gamma_linear_stark(double n_up, double n_lo, double n_e) {  // broadening.py:193-229
    double a1 = (n_up - n_lo < 1.5) ? 0.642 : 1.0;
    return 0.60 * a1 * (n_up * n_up - n_lo * n_lo) * pow(n_e, 2.0 / 3.0);
}
SD_HD double gamma_quadratic_stark(double z_eff, double n_up, double n_lo, double n_e, double T) {  // :281-344
    const double eps0 = 1.0 / (4.0 * PI);
    double pref = (E_ESU * E_ESU * A0_CGS * A0_CGS * A0_CGS) / (36.0 * H_CGS * eps0 * z_eff * z_eff * z_eff * z_eff);
    double t1 = n_up * ((5.0 * n_up * n_up) + 1.0);
    double t2 = n_lo * ((5.0 * n_lo * n_lo) + 1.0);
    double c4 = pref * (t1 * t1 - t2 * t2);
    return 1e19 * KB_CGS * n_e * pow(c4, 2.0 / 3.0) * pow(T, 1.0 / 6.0);
}
SD_HD double gamma_van_der_waals(double z_eff, double n_up, double n_lo, double T, double n_H) {  // :420-472
    double u2 = n_up * n_up, l2 = n_lo * n_lo;
    double c6 = 6.46e-34 * ((5.0 * u2 * u2 + u2) - (5.0 * l2 * l2 + l2)) / (2.0 * z_eff * z_eff);
    return 17.0 * pow(8.0 * KB_CGS * T / (PI * MP_CGS), 0.3) * pow(c6, 0.4) * n_H;
}
SD_HD double vald_stark(double n_e, double stark, double T) {  // broadening.py:880-890
    double g = n_e * pow(10.0, stark) * pow(T / 1e4, 1.0 / 6.0);
    return (n_e * stark >= 0) ? 0.0 : g;
}
// calc_vald_vdW (broadening.py:893-1006) for n_H = 1; caller multiplies by the hydrogen density.
SD_HD double vald_vdw_unit(double vdw, double z_eff, double n_up, double n_lo, double T, double mass) {
    if (vdw < 0) return pow(10.0, vdw) * pow(T / 1e4, 0.38);
    if (vdw == 0.0) return 0.0;
    if (vdw < 20) return gamma_van_der_waals(z_eff, n_up, n_lo, T, 1.0) * vdw;
    if (!(vdw >= 20)) return 0.0;  // NaN code: no mask matches in the reference -> stays 0
    double vi = (double)(long long)vdw;
    double sigma = vi * A0_CGS * A0_CGS;
    double alpha = vdw - vi;
    double inv_mu = 1.0 / (1.008 * AMU_CGS) + (1.0 / mass);
    double vbar = sqrt(8.0 * KB_CGS * T / PI * inv_mu);
    return 2.0 * pow(4.0 / PI, alpha / 2.0) * tgamma((4.0 - alpha) / 2.0) * 1e6 * sigma * pow(vbar / 1e6, 1.0 - alpha);
}

// ---------------------------------------------------------------------------- line window
// opacities_solvers/base.py:556-575 incl. the int() overflow quirk (SURVEY 8a K2 ii/iii).
// idx = number of grid points with nu >= nu_line; returns [lo, hi) on the global grid.
SD_HD void line_window(long long idx, long long N, double gamma, double dw, double alpha, double d_nu,
                       long long &lo, long long &hi) {
    double broad = ((gamma + dw) * alpha) / d_nu * 20.0;
    double forced = (broad > 10.0) ? broad : 10.0;
    long long hw = (forced < 9.2233720368547758e18) ? (long long)forced : INT64_MIN;
    long long a = (long long)((unsigned long long)idx - (unsigned long long)hw);
    long long b = (long long)((unsigned long long)idx + (unsigned long long)hw);
    long long l = a > 0 ? a : 0;
    long long h = b < N ? b : N;
    if (h < 0) {
        h += N;
        if (h < 0) h = 0;
    }
    if (h < l) h = l;
    lo = l;
    hi = h;
}

// ---------------------------------------------------------------------------- formal solver
SD_HD double planck(double nu, double T) {  // source_functions/blackbody.py:31-35
    double pre = (2.0 * H_CGS * nu * nu * nu) / (C_CGS * C_CGS);
    return pre / (exp((H_CGS * nu) / (KB_CGS * T)) - 1.0);
}
SD_HD void rt_weights(double tau, double &w0, double &w1, double &w2) {  // radiation_field_solvers/base.py:6-47
    if (tau < 5e-4) {
        w0 = tau * (1.0 - tau / 2.0);
        w1 = tau * tau * (0.5 - tau / 3.0);
        w2 = tau * tau * tau * (1.0 / 3.0 - tau / 4.0);
    } else if (tau < 50.0) {
        double e = exp(-tau);
        w0 = 1.0 - e;
        w1 = w0 - tau * e;
        w2 = 2.0 * w1 - tau * tau * e;
    } else {
        w0 = 1.0;
        w1 = 1.0;
        w2 = 2.0;
    }
}

}  // namespace sdm
