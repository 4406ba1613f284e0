// k4_raytrace.cu -- K4: formal solution of radiative transfer for ALL angles in one pass + flux quadrature.
//
// Reference: stardis/radiation_field/radiation_field_solvers/base.py:271-346 (raytrace: serial Python loop
// over angles, each calling single_theta_trace_parallel :85-268, which materialises ~8 (D,N) temporaries:
// log-mean opacity, tau, Planck source, three weight arrays) and :6-47 (calc_weights_parallel);
// source_functions/blackbody.py:11-35.
//
// B200 design: one thread owns one frequency.  The depth recurrence runs with depth as the OUTER loop and the
// angle as the INNER (fully unrolled) loop, so the per-angle intensities I[theta] stay in registers, the
// depth-only quantities (Planck source S_k, sqrt(alpha_k)) are computed once per cell instead of once per
// angle, and F_nu[k] = sum_theta w_theta I_theta[k] is written coalesced as soon as depth k is finished.
// alpha is read once (twice with inward rays), F_nu written once: 16 B per (depth, nu) cell.
//
// Geometric mean of the opacity (base.py:121): exp((log a + log b)/2) is evaluated as sqrt(a) * sqrt(b)
// (same NaN/0/inf behaviour; agrees to a few ulp, far inside the 1e-6 tolerance on F_nu).
#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

constexpr int MAX_TH = 20;  // angles per launch (kept in registers); more angles -> several launches

struct RayArgs {
    int D;
    int64_t N, p0, p1;
    const double *nus;     // [N]
    const double *T;       // [D]
    const double *alpha;   // (D, W)
    const double *ds;      // (G, n_theta_total) row-major
    const double *wts;     // [n_theta_total]
    int th0, th_total;     // this launch handles angles [th0, th0 + TH)
    int inward;
    int accumulate;        // add to the existing F
    double scale;          // applied to the final F (last launch only)
    double *F;             // (D, W)
    double *I_nus;         // (D, W, th_total) or nullptr
};

// One step of the short-characteristics recurrence (base.py:209-249 / :151-198) with the five divisions of the
// reference replaced by two reciprocals, one of which (1/tau of the gap just crossed) is carried over to the next step:
//   second + third = [ dA r_n (w1 tau_k - w2) - dB r_k (w1 tau_n + w2) ] / (tau_k + tau_n),
//   dA = S_mid - S_far, dB = S_mid - S_near, r = 1/tau.
// Outward: tau_k = this gap, tau_n = next gap, S_mid = S_{k+1}, S_far = S_{k+2}, S_near = S_k; the inward sweep uses
// the same expression with mirrored arguments.  A zero tau_n gives NaN like the reference's 0-division does.
__device__ __forceinline__ double sc_step(double I_prev, double tau_k, double r_k, double tau_n, double r_n, double S_mid,
                                          double dA, double dB) {
    double w0, w1, w2;
    sdm::rt_weights(tau_k, w0, w1, w2);
    const double rs = sdm::rcp_fast(tau_k + tau_n);
    const double u = fma(w1, tau_k, -w2), v = fma(w1, tau_n, w2);
    const double corr = fma(dA * r_n, u, -(dB * r_k) * v) * rs;
    return fma(1.0 - w0, I_prev, fma(w0, S_mid, corr));
}

template <int TH>
__global__ void __launch_bounds__(128) k_raytrace(RayArgs a) {
    extern __shared__ double smem[];  // ds for this launch's angles: [G][TH], then T[D], then w[TH]
    const int D = a.D, G = a.D - 1;
    double *s_ds = smem;
    double *s_T = smem + (size_t)G * TH;
    double *s_w = s_T + D;
    for (int k = threadIdx.x; k < G * TH; k += blockDim.x) {
        int g = k / TH, t = k - g * TH;
        s_ds[k] = a.ds[(size_t)g * a.th_total + a.th0 + t];
    }
    for (int k = threadIdx.x; k < D; k += blockDim.x) s_T[k] = a.T[k];
    for (int k = threadIdx.x; k < TH; k += blockDim.x) s_w[k] = a.wts[a.th0 + k];
    __syncthreads();

    const int64_t W = a.p1 - a.p0;
    const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (col >= W) return;
    const double nu = a.nus[a.p0 + col];
    const double *al = a.alpha + col;
    const bool last = a.scale != 0.0;  // scale == 0 marks "not the last launch"
    const double scale = last ? a.scale : 1.0;

    double I[TH], tau_c[TH], r_c[TH];  // intensity, optical depth of the current gap and its reciprocal, per angle
#pragma unroll
    for (int t = 0; t < TH; t++) I[t] = 0.0;

    auto emit = [&](int k) {
        double f = 0.0;
#pragma unroll
        for (int t = 0; t < TH; t++) f += I[t] * s_w[t];  // angle order as base.py:324-338
        size_t o = (size_t)k * W + col;
        if (a.accumulate) f += a.F[o];
        a.F[o] = f * scale;
        if (a.I_nus) {
#pragma unroll
            for (int t = 0; t < TH; t++) a.I_nus[o * a.th_total + a.th0 + t] = I[t];
        }
    };

    if (a.inward) {
        // base.py:141-198: from the surface (I = 0) down to the deepest point.  gap index k-1 == -1 wraps to
        // the LAST gap / LAST depth point (numpy negative indexing), reproduced here.
        const double sa_last = sqrt(al[(size_t)(D - 1) * W]), sa_last2 = sqrt(al[(size_t)(D - 2) * W]);
        const double mean_wrap = sa_last * sa_last2;  // mean opacity of gap G-1
        const double S_wrap = sdm::planck(nu, s_T[D - 1]);
        double sa_k = sa_last2;
        double S_kp1 = S_wrap, S_k = sdm::planck(nu, s_T[D - 2]);
#pragma unroll
        for (int t = 0; t < TH; t++) {
            tau_c[t] = mean_wrap * s_ds[(G - 1) * TH + t];
            r_c[t] = sdm::rcp_fast(tau_c[t]);
        }
        for (int k = G - 1; k >= 0; k--) {
            double sa_km1 = 0.0, S_km1, mean_km;
            if (k >= 1) {
                sa_km1 = sqrt(al[(size_t)(k - 1) * W]);
                S_km1 = sdm::planck(nu, s_T[k - 1]);
                mean_km = sa_k * sa_km1;
            } else {
                S_km1 = S_wrap;
                mean_km = mean_wrap;
            }
            const int km = (k >= 1) ? k - 1 : G - 1;
            const double dA = S_k - S_km1, dB = S_k - S_kp1;
#pragma unroll
            for (int t = 0; t < TH; t++) {
                const double tau_k = tau_c[t], r_k = r_c[t];
                const double tau_m = mean_km * s_ds[km * TH + t];
                const double r_m = sdm::rcp_fast(tau_m);
                if (!(tau_k == 0.0 || tau_m == 0.0)) I[t] = sc_step(I[t], tau_k, r_k, tau_m, r_m, S_k, dA, dB);
                tau_c[t] = tau_m;
                r_c[t] = r_m;
            }
            sa_k = sa_km1;
            S_kp1 = S_k; S_k = S_km1;
        }
    }
    emit(0);

    // outward sweep, base.py:200-266
    double sa1 = sqrt(al[(size_t)W]);
    double S0 = sdm::planck(nu, s_T[0]), S1 = sdm::planck(nu, s_T[1]);
    {
        const double mean0 = sqrt(al[0]) * sa1;
#pragma unroll
        for (int t = 0; t < TH; t++) {
            tau_c[t] = mean0 * s_ds[t];
            r_c[t] = sdm::rcp_fast(tau_c[t]);
        }
    }
    // the opacity of depth k + 3 is requested one iteration before its square root is needed: at 12 warps per SM the
    // latency of this load sat on the critical path of every depth step (ncu round 3: 15 % of the samples)
    double al_next = al[(size_t)(2 < D ? 2 : D - 1) * W];
    for (int k = 0; k < G - 1; k++) {
        const double al_cur = al_next;
        al_next = al[(size_t)(k + 3 < D ? k + 3 : D - 1) * W];
        const double sa2 = sqrt(al_cur);
        const double S2 = sdm::planck(nu, s_T[k + 2]);
        const double mean1 = sa1 * sa2;
        const double dA = S1 - S2, dB = S1 - S0;
#pragma unroll
        for (int t = 0; t < TH; t++) {
            const double tau_k = tau_c[t], r_k = r_c[t];
            const double tau_n = mean1 * s_ds[(k + 1) * TH + t];
            const double r_n = sdm::rcp_fast(tau_n);
            if (tau_k != 0.0) I[t] = sc_step(I[t], tau_k, r_k, tau_n, r_n, S1, dA, dB);
            tau_c[t] = tau_n;
            r_c[t] = r_n;
        }
        emit(k + 1);
        sa1 = sa2;
        S0 = S1; S1 = S2;
    }
    {  // final jump, base.py:253-266 (after the loop: S0 = S_{D-2}, S1 = S_{D-1}, tau_c = gap G-1)
#pragma unroll
        for (int t = 0; t < TH; t++) {
            const double tau = tau_c[t];
            if (tau != 0.0) {
                double w0, w1, w2;
                sdm::rt_weights(tau, w0, w1, w2);
                const double third = w2 * (S0 - S1) * (r_c[t] * r_c[t]);
                I[t] = (1.0 - w0) * I[t] + w0 * S1 + third;
            }
        }
        emit(D - 1);
    }
}

template <int TH>
int launch_th(sd_ctx *c, RayArgs &a) {
    size_t smem = sizeof(double) * ((size_t)(a.D - 1) * TH + a.D + TH);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_raytrace<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t W = a.p1 - a.p0;
    k_raytrace<TH><<<(unsigned)((W + 127) / 128), 128, smem, c->stream>>>(a);
    return sd_launch_check(c, "k_raytrace");
}

int launch_dyn(sd_ctx *c, RayArgs &a, int th) {
    switch (th) {
#define CASE(n) case n: return launch_th<n>(c, a);
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
        CASE(11) CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18) CASE(19) CASE(20)
#undef CASE
    }
    return sd_fail(c, SD_ERR_ARG, "bad angle chunk %d", th);
}

}  // namespace

int sd_k4_raytrace(sd_ctx *c, int n_theta, const double *ray_ds, const double *weights, int inward, double scale,
                   int track) {
    const int D = c->D, G = D - 1;
    const int64_t W = c->W();
    SD_CHECK(c, scale != 0.0, SD_ERR_ARG, "sd_raytrace: scale must be non-zero");
    SD_CHECK(c, sizeof(double) * ((size_t)G * MAX_TH + D + MAX_TH) <= 200 * 1024, SD_ERR_ARG,
             "sd_raytrace: too many depth points for the shared-memory path table");
    size_t nds = (size_t)G * n_theta;
    SD_TRY(sd_ensure(c, c->ray_small, sizeof(double) * (nds + n_theta)));
    double *d_ds = c->ray_small.as<double>(), *d_w = d_ds + nds;
    SD_CUDA(c, cudaMemcpyAsync(d_ds, ray_ds, sizeof(double) * nds, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(d_w, weights, sizeof(double) * n_theta, cudaMemcpyDefault, c->stream));
    SD_TRY(sd_ensure(c, c->F, sizeof(double) * D * W));
    if (track) SD_TRY(sd_ensure(c, c->I_nus, sizeof(double) * D * W * n_theta));
    sd_phase_begin(c, SD_PH_K4);  // after the small descriptor copies: the kernel launches only
    RayArgs a{};
    a.D = D; a.N = c->N; a.p0 = c->p0; a.p1 = c->p1;
    a.nus = c->nus.as<double>(); a.T = c->T.as<double>(); a.alpha = c->total.as<double>();
    a.ds = d_ds; a.wts = d_w; a.th_total = n_theta; a.inward = inward;
    a.F = c->F.as<double>(); a.I_nus = track ? c->I_nus.as<double>() : nullptr;
    for (int th0 = 0; th0 < n_theta; th0 += MAX_TH) {
        int th = (n_theta - th0 < MAX_TH) ? n_theta - th0 : MAX_TH;
        a.th0 = th0;
        a.accumulate = th0 > 0;
        a.scale = (th0 + th >= n_theta) ? scale : 0.0;
        SD_TRY(launch_dyn(c, a, th));
    }
    return SD_OK;
}
