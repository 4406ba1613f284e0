// Host-side harness that compiles stardis_b200/csrc/sd_math.cuh (the scalar arithmetic shared by every
// CUDA kernel) with the HOST compiler, so that the formulas are checked against the reference's golden
// vectors in the CPU-only test suite.  Test infrastructure: not part of the product library.
#include "../stardis_b200/csrc/sd_math.cuh"
extern "C" {
void hm_humlicek(long n, const double *x, const double *y, double *re, double *wr, double *wi, int *region) {
    for (long i = 0; i < n; i++) {
        re[i] = sdm::humlicek_re(x[i], y[i]);
        sdm::humlicek_complex(x[i], y[i], wr[i], wi[i]);
        region[i] = sdm::humlicek_region(x[i], y[i]);
    }
}
void hm_voigt(long n, const double *dnu, const double *dw, const double *g, double *phi) {
    for (long i = 0; i < n; i++) phi[i] = sdm::voigt_profile(dnu[i], dw[i], g[i]);
}
void hm_region1(long n, const double *x, const double *y, double *out) {
    for (long i = 0; i < n; i++) out[i] = sdm::region1_re(x[i] * x[i], y[i]);
}
void hm_broadening(long n, const double *zeff, const double *nu, const double *nl, const double *ne, const double *T,
                   const double *nH, double *ls, double *qs, double *vdw) {
    for (long i = 0; i < n; i++) {
        ls[i] = sdm::gamma_linear_stark(nu[i], nl[i], ne[i]);
        qs[i] = sdm::gamma_quadratic_stark(zeff[i], nu[i], nl[i], ne[i], T[i]);
        vdw[i] = sdm::gamma_van_der_waals(zeff[i], nu[i], nl[i], T[i], nH[i]);
    }
}
void hm_neff_doppler(long n, const double *zeff, const double *eion, const double *elev, const double *nuline,
                     const double *T, const double *mass, double vmic, double *neff, double *dw) {
    for (long i = 0; i < n; i++) {
        neff[i] = sdm::n_effective(zeff[i], eion[i], elev[i]);
        dw[i] = sdm::doppler_width(nuline[i], T[i], mass[i], vmic);
    }
}
void hm_vald(long n, const double *vdw, const double *stark, const double *zeff, const double *nu, const double *nl,
             const double *T, const double *mass, const double *ne, double *g_vdw, double *g_stark) {
    for (long i = 0; i < n; i++) {
        g_vdw[i] = sdm::vald_vdw_unit(vdw[i], zeff[i], nu[i], nl[i], T[i], mass[i]);
        g_stark[i] = sdm::vald_stark(ne[i], stark[i], T[i]);
    }
}
void hm_window(long n, const long long *idx, long long N, const double *g, const double *dw, const double *a, double d_nu,
               long long *lo, long long *hi) {
    for (long i = 0; i < n; i++) sdm::line_window(idx[i], N, g[i], dw[i], a[i], d_nu, lo[i], hi[i]);
}
void hm_planck_weights(long n, const double *nu, const double *T, const double *tau, double *B, double *w0, double *w1,
                       double *w2) {
    for (long i = 0; i < n; i++) {
        B[i] = sdm::planck(nu[i], T[i]);
        sdm::rt_weights(tau[i], w0[i], w1[i], w2[i]);
    }
}
}
