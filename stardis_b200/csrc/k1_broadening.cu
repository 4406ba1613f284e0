// k1_broadening.cu -- K1: collisional broadening and Doppler widths per (line, depth), and the K2
// preparation pass (global window of every (line, depth) pair, half-width classes, 64-byte line records).
//
// Reference: stardis/radiation_field/opacities/opacities_solvers/broadening.py:550-656 (calc_gamma),
// :659-732 (calculate_broadening: z_eff = ion_number + 1, linear Stark for hydrogen only), :1009-1085
// (calc_vald_gamma), :32-71 (calc_doppler_width); opacities_solvers/base.py:522-575 (window rule).
//
// K1 layout: every pow() of the formulae depends either on the line only (effective quantum numbers, C4^(2/3), C6^0.4,
// 10^stark ...) or on the depth only (n_e^(2/3), T^(1/6), (8kT/pi m_p)^0.3 ...), never on both (the ABO van der Waals
// code of VALD lists excepted).  k_line_pre / k_depth_pre evaluate them once per line / per depth; k_broadening then
// multiplies the factors in the reference's order of association per (line, depth) pair -- the same numbers as the
// unhoisted formula, bit for bit -- and is bound by its two (L, D) stores instead of by ~6 FP64 pow() per pair.
//
// Roofline: HBM.  Per (line, depth): K1 writes gamma 8 B + doppler width 8 B; the preparation pass reads those and
// alpha_line (24 B) and writes window 32 B + class 1 B, and -- only for pairs whose window meets this context's extended
// pixel range -- the 64-byte record and up to two 8-byte window-edge keys.
#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

constexpr int LP_N = 8;    // doubles per line in line_pre
constexpr int DP_N = 10;   // doubles per depth in depth_pre
enum { LP_LS = 0, LP_QS = 1, LP_VDW = 2, LP_RAD = 3, LP_NUC = 4, LP_MASS = 5, LP_WAALS = 6, LP_STARK = 7 };
enum { DP_NE23 = 0, DP_AQS = 1, DP_T16 = 2, DP_P17 = 3, DP_NH = 4, DP_2KT = 5, DP_NE = 6, DP_T16V = 7, DP_T038 = 8, DP_T = 9 };

// per-line factors.  non-VALD (broadening.py:611-654): LS = 0.60 a1 (n_u^2 - n_l^2) [hydrogen only], QS = C4^(2/3),
// VDW = C6^0.4.  VALD (:1039-1085): LS as above, QS = 10^stark, VDW = 10^waals (code < 0) or C6^0.4 (0 < code < 20).
__global__ void __launch_bounds__(256) k_line_pre(int64_t L, uint32_t flags, const double *__restrict__ nu,
                                                  const int64_t *__restrict__ Z, const int64_t *__restrict__ ion,
                                                  const double *__restrict__ e_ion, const double *__restrict__ e_up,
                                                  const double *__restrict__ e_lo, const double *__restrict__ A_ul,
                                                  const double *__restrict__ mass, const double *__restrict__ stark,
                                                  const double *__restrict__ waals, double *__restrict__ pre) {
    const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= L) return;
    const double zeff = (double)(ion[l] + 1);
    const double n_up = sdm::n_effective(zeff, e_ion[l], e_up[l]);
    const double n_lo = sdm::n_effective(zeff, e_ion[l], e_lo[l]);
    double *o = pre + l * LP_N;
    // linear Stark, hydrogen lines only (broadening.py:614-620): 0.60 * a1 * (n_u^2 - n_l^2) [* n_e^(2/3)]
    const bool ls = (flags & SD_LINEAR_STARK) && Z[l] == 1;
    const double a1 = (n_up - n_lo < 1.5) ? 0.642 : 1.0;
    o[LP_LS] = ls ? 0.60 * a1 * (n_up * n_up - n_lo * n_lo) : 0.0;
    // C6^0.4 (broadening.py:420-472)
    const double u2 = n_up * n_up, l2 = n_lo * n_lo;
    const double c6 = 6.46e-34 * ((5.0 * u2 * u2 + u2) - (5.0 * l2 * l2 + l2)) / (2.0 * zeff * zeff);
    const double c6_04 = pow(c6, 0.4);
    if (flags & SD_VALD) {
        const double w = waals[l];
        o[LP_QS] = pow(10.0, stark[l]);
        o[LP_VDW] = (w < 0) ? pow(10.0, w) : c6_04;
        o[LP_WAALS] = w;
        o[LP_STARK] = stark[l];
    } else {
        // C4^(2/3) (broadening.py:281-344)
        const double eps0 = 1.0 / (4.0 * sdm::PI);
        const double pref = (sdm::E_ESU * sdm::E_ESU * sdm::A0_CGS * sdm::A0_CGS * sdm::A0_CGS) /
                            (36.0 * sdm::H_CGS * eps0 * zeff * zeff * zeff * zeff);
        const double t1 = n_up * ((5.0 * n_up * n_up) + 1.0);
        const double t2 = n_lo * ((5.0 * n_lo * n_lo) + 1.0);
        const double c4 = pref * (t1 * t1 - t2 * t2);
        o[LP_QS] = pow(c4, 2.0 / 3.0);
        o[LP_VDW] = c6_04;
        o[LP_WAALS] = 0.0;
        o[LP_STARK] = 0.0;
    }
    o[LP_RAD] = A_ul[l];
    o[LP_NUC] = nu[l] / sdm::C_CGS;
    o[LP_MASS] = mass[l];
}

__global__ void k_depth_pre(int D, const double *__restrict__ T, const double *__restrict__ ne, const double *__restrict__ nH,
                            double *__restrict__ pre) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double *o = pre + d * DP_N;
    const double Td = T[d], ned = ne[d];
    o[DP_NE23] = pow(ned, 2.0 / 3.0);
    o[DP_AQS] = 1e19 * sdm::KB_CGS * ned;
    o[DP_T16] = pow(Td, 1.0 / 6.0);
    o[DP_P17] = 17.0 * pow(8.0 * sdm::KB_CGS * Td / (sdm::PI * sdm::MP_CGS), 0.3);
    o[DP_NH] = nH[d];
    o[DP_2KT] = 2.0 * sdm::KB_CGS * Td;
    o[DP_NE] = ned;
    o[DP_T16V] = pow(Td / 1e4, 1.0 / 6.0);
    o[DP_T038] = pow(Td / 1e4, 0.38);
    o[DP_T] = Td;
}

// One thread per (line, depth), depth fastest: the 64-byte per-line block is a broadcast within a line, the per-depth
// block comes from shared memory; two coalesced (L, D) stores.
__global__ void __launch_bounds__(256) k_broadening(int64_t L, int D, const double *__restrict__ line_pre,
                                                    const double *__restrict__ depth_pre, double vmic, uint32_t flags,
                                                    double *__restrict__ gammas, double *__restrict__ dws) {
    extern __shared__ double s_dp[];  // [D][DP_N]
    for (int k = threadIdx.x; k < D * DP_N; k += blockDim.x) s_dp[k] = depth_pre[k];
    __syncthreads();
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // (l, d), d fastest
    if (g >= L * D) return;
    const int64_t l = g / D;
    const int d = (int)(g - l * D);
    const double2 *lp2 = reinterpret_cast<const double2 *>(line_pre + l * LP_N);
    const double2 p01 = __ldg(lp2), p23 = __ldg(lp2 + 1), p45 = __ldg(lp2 + 2), p67 = __ldg(lp2 + 3);
    const double *dp = s_dp + d * DP_N;
    double gam;
    if (flags & SD_VALD) {  // broadening.py:1039-1085
        gam = 0.0;
        if (flags & SD_RADIATION) gam += p23.y;
        if (flags & SD_LINEAR_STARK) gam += p01.x * dp[DP_NE23];  // 0 for non-hydrogen lines (finite densities)
        if (flags & SD_QUADRATIC_STARK) {  // calc_vald_stark_gamma :880-890
            const double gq = dp[DP_NE] * p01.y * dp[DP_T16V];
            gam += (dp[DP_NE] * p67.y >= 0) ? 0.0 : gq;
        }
        if (flags & SD_VAN_DER_WAALS) {  // calc_vald_vdW :893-1006 (n_H = 1) * n_H
            const double w = p67.x;
            double unit;
            if (w < 0) unit = p23.x * dp[DP_T038];
            else if (w == 0.0) unit = 0.0;
            else if (w < 20) unit = dp[DP_P17] * p23.x * 1.0 * w;
            else if (!(w >= 20)) unit = 0.0;
            else unit = sdm::vald_vdw_unit(w, 1.0, 1.0, 1.0, dp[DP_T], p45.y);  // ABO: T and the line mixed in one pow()
            gam += unit * dp[DP_NH];
        }
        gam /= 2.0;
    } else {  // broadening.py:611-654
        double g_ls = 0.0, g_qs = 0.0, g_vdw = 0.0, g_rad = 0.0;
        if (flags & SD_LINEAR_STARK) g_ls = p01.x * dp[DP_NE23];
        if (flags & SD_QUADRATIC_STARK) g_qs = dp[DP_AQS] * p01.y * dp[DP_T16];
        if (flags & SD_VAN_DER_WAALS) g_vdw = dp[DP_P17] * p23.x * dp[DP_NH];
        if (flags & SD_RADIATION) g_rad = p23.y;
        gam = g_ls + g_qs + g_vdw + g_rad;
    }
    gammas[g] = gam;
    dws[g] = p45.x * sqrt(dp[DP_2KT] / p45.y + vmic * vmic);  // calc_doppler_width :32-66
}

// d_nu = -max(diff(nus))  (opacities_solvers/base.py:524-526); one block.
__global__ void __launch_bounds__(1024) k_dnu(int64_t N, const double *__restrict__ nus, double *__restrict__ out) {
    __shared__ double sm[32];
    double m = -INFINITY;
    for (int64_t i = 1 + threadIdx.x; i < N; i += blockDim.x) m = fmax(m, nus[i] - nus[i - 1]);
    // NaN-free grids assumed (a NaN frequency is a caller error); fmax ignores NaN like np.max would not.
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : -INFINITY;
        for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) out[0] = -m;
    }
}

// idx[l] = number of grid points with nu >= nu_line on the descending grid
//        = N - searchsorted(nus[::-1], nu_line, 'left')   (opacities_solvers/base.py:556-558)
__global__ void __launch_bounds__(256) k_line_idx(int64_t L, int64_t N, const double *__restrict__ nus,
                                                  const double *__restrict__ line_nu, int *__restrict__ idx) {
    int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= L) return;
    double v = line_nu[l];
    int64_t lo = 0, hi = N;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        if (nus[mid] >= v) lo = mid + 1; else hi = mid;
    }
    idx[l] = (int)lo;
}

// Centre frequency, half-width and moment scale [Hz] of every GLOBAL pixel tile [t*T, min((t+1)*T, N)) (far-field
// expansion point).  The moment scale is the half-width, except for a ragged last tile, which takes its neighbour's.
__global__ void __launch_bounds__(256) k_tile_geometry(int64_t N, int tile, int n_tiles, const double *__restrict__ nus,
                                                       double *__restrict__ geom) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int64_t a = (int64_t)t * tile, b = a + tile < N ? a + tile : N;
    double hi = nus[a], lo = nus[b - 1];
    geom[3 * t] = 0.5 * (hi + lo);
    geom[3 * t + 1] = 0.5 * (hi - lo);
    double sc = 0.5 * (hi - lo);
    if (b - a < tile && t > 0) sc = 0.5 * (nus[a - tile] - nus[a - 1]);
    geom[3 * t + 2] = sc;
}

// Which hierarchy levels are usable on this grid: the far criterion is an index distance (two tiles), so the tile
// half-widths must vary smoothly (then the nearest far tile is the worst case of every convergence test made per
// pair by k_build_records).  Level k is active when it and all lower levels have strictly descending frequencies and
// neighbouring full tiles differ by less than a factor 1.6 in width.  info[0] = number of active levels.
__global__ void __launch_bounds__(256) k_level_check(FarGeom fg, int64_t N, int *__restrict__ info) {
    __shared__ int s_bad[SD_FAR_LEVELS];
    if (threadIdx.x < SD_FAR_LEVELS) s_bad[threadIdx.x] = 0;
    __syncthreads();
    for (int k = 0; k < SD_FAR_LEVELS; k++) {
        const int nt = fg.n_tiles[k];
        const int n_full = (int)(N / fg.tile[k]);  // tiles [0, n_full) hold a full set of pixels
        bool bad = false;
        for (int t = threadIdx.x; t < nt; t += blockDim.x) {
            const double h = fg.geom[k][3 * t + 1];
            const bool single = (t == nt - 1) && (N - (int64_t)t * fg.tile[k] < 2);
            if (!(h > 0.0) && !single) bad = true;
            if (t + 1 < n_full) {
                const double r = fg.geom[k][3 * (t + 1) + 1] / h;
                if (!(r > 0.625 && r < 1.6)) bad = true;
            }
        }
        if (bad) s_bad[k] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        while (n < SD_FAR_LEVELS && !s_bad[n]) n++;
        info[0] = n;
    }
}

__device__ __forceinline__ int hw_class(long long hw) {
    // class 0: hw <= 64; class k (1..5): hw <= 64 * 4^k; class 6: everything wider (walked whole by the line kernel).
    // Classes SD_FC0 + m are reserved for the far-capable pairs of the far-field scheme.
    if (hw <= SD_CLS0_HW) return 0;
    int k = 1;
    long long lim = (long long)SD_CLS0_HW * 4;
    while (k < SD_FC0 - 1 && hw > lim) { k++; lim *= 4; }
    return k;
}

// Can the pair (nu_l, dw, y; centre pixel cpix) be expanded at hierarchy level k?  Everything two or more tiles away
// from the centre tile s must (i) be certainly in Humlicek region I (x^2 > thr at the nearest such pixel on either
// side), (ii) see the multipole series of the two poles about the centre of s converge with ratio <= SD_FAR_RHO
// (|pole - c_s| against the distance of the nearest far pixel from c_s) and (iii) see the Taylor series about its own
// centre converge with that ratio (half-width of the nearest far tile against its distance from the nearer pole).
// The nearest far tile is the worst case on a smooth grid (k_level_check).  Written as positive conditions: NaN
// parameters never qualify.  A side without far tiles imposes nothing.
__device__ __forceinline__ bool level_ok(const FarGeom &fg, int k, int64_t N, const double *__restrict__ nus, int cpix, double nu_l,
                                         double dw, double inv_dw, double g, double thr) {
    const int T = fg.tile[k], nt = fg.n_tiles[k];
    const int s = cpix >> fg.tile_shift[k];
    const double *__restrict__ gm = fg.geom[k];
    const double c_s = gm[3 * s];
    const double adw = 0.7071067811865476 * dw;
    const double off = fabs(nu_l - c_s) + adw;
    const double pole2 = fma(off, off, g * g);
    const double reach = (1.0 + SD_FAR_OVERHANG) * gm[3 * s + 2];  // k_m2l shortens its sums on this assumption
    bool ok = (dw > 0.0) && (g >= 0.0) && (pole2 <= reach * reach);
    if (s - 2 >= 0) {
        const double nu_e = nus[(int64_t)(s - 1) * T - 1];      // last pixel of tile s - 2 (frequencies descend)
        const double x = (nu_e - nu_l) * inv_dw, de = nu_e - c_s;
        const double c_t = gm[3 * (s - 2)], h_t = gm[3 * (s - 2) + 1];
        ok = ok && (x * x > thr) && (pole2 <= SD_FAR_RHO * SD_FAR_RHO * de * de) && (h_t <= SD_FAR_RHO * (c_t - nu_l - adw));
    }
    if (s + 2 < nt) {
        const double nu_e = nus[(int64_t)(s + 2) * T];          // first pixel of tile s + 2
        const double x = (nu_e - nu_l) * inv_dw, de = nu_e - c_s;
        const double c_t = gm[3 * (s + 2)], h_t = gm[3 * (s + 2) + 1];
        ok = ok && (x * x > thr) && (pole2 <= SD_FAR_RHO * SD_FAR_RHO * de * de) && (h_t <= SD_FAR_RHO * (nu_l - adw - c_t));
    }
    return ok;
}

// One thread per (line, depth).  A block takes a tile of 32 lines x 8 depth points: the (L,D) inputs are read as 64-byte
// segments (8 depths of a line), the depth-major outputs are staged in shared memory and written as contiguous rows --
// 32 records = 2 KB, 32 window records = 512 B, 32 classes per depth point (one thread per pair with the depth fastest
// wrote 64-byte records 19 MB apart and half-sector window records).
__global__ void __launch_bounds__(256) k_build_records(int64_t L, int D, int64_t N, const double *__restrict__ nus,
                                                       const double *__restrict__ line_nu,
                                                       const int *__restrict__ line_idx, const double *__restrict__ gammas,
                                                       int gamma_cols, const double *__restrict__ dws,
                                                       const double *__restrict__ alpha, const double *__restrict__ d_nu_p,
                                                       LineRec *__restrict__ rec, PairWin *__restrict__ win,
                                                       uint8_t *__restrict__ win_cls,
                                                       FarGeom fg, unsigned long long *__restrict__ stats) {
    constexpr int TL = 32, TD = 8;                       // tile: lines x depth points (TL * TD = blockDim.x)
    __shared__ __align__(16) LineRec s_rec[TD][TL];
    __shared__ __align__(16) PairWin s_win[TD][TL];
    __shared__ uint8_t s_wcls[TD][TL];
    const int n_dt = (D + TD - 1) / TD;
    const unsigned l0 = (blockIdx.x / (unsigned)n_dt) * TL;
    const int d0 = (int)(blockIdx.x % (unsigned)n_dt) * TD;
    const int tl = threadIdx.x / TD, td = threadIdx.x % TD;
    const unsigned l = l0 + tl;
    const int d = d0 + td;
    const unsigned g = l * (unsigned)D + d;                  // L * D < 2^31 (checked by sd_set_lines)
    bool active = (l < (unsigned)L) && (d < D);
    unsigned nonempty = 0, wide = 0, zero_dw = 0;
    bool e_lo = false, e_hi = false;
    unsigned long long key_lo = 0, key_hi = 0;
    if (active) {
        double gam = gamma_cols > 1 ? gammas[g] : gammas[l];  // base.py:547-551
        double dw = dws[g];
        double a = alpha[g];
        double d_nu = d_nu_p[0];
        const int idx = line_idx[l];
        long long lo, hi;
        sdm::line_window(idx, N, gam, dw, a, d_nu, lo, hi);
        // half-width as the reference computes it, for the class only
        double broad = ((gam + dw) * a) / d_nu * 20.0;
        double forced = (broad > 10.0) ? broad : 10.0;
        long long hw = (forced < 4.0e18) ? (long long)forced : (long long)4e18;
        int cls = hw_class(hw);
        double y = (gam / sdm::SQRT_PI_PI) / dw;
        LineRec r;
        r.nu = line_nu[l];
        r.inv_dw = 1.0 / dw;
        r.dw = dw;
        r.y = y;
        r.K = a / (sdm::SQRT_PI * dw);
        double m = 15.0 * (1.0 + 1e-9) - y;       // |x| > m  =>  |x| + y > 15 with a safe margin
        r.thr = (m > 0.0) ? m * m : ((m <= 0.0) ? -1.0 : m);  // NaN y -> NaN thr -> never "far"
        if (!(r.inv_dw > 0.0) || !(r.inv_dw < 1e300)) r.thr = NAN;  // dw <= 0, inf or NaN: exact path only
        r.pad0 = r.pad1 = 0.0;
        s_rec[td][tl] = r;
        PairWin pw;
        pw.lo = (int)lo;
        pw.hi = (int)hi;
        const int cpix = idx < N ? idx : (int)(N - 1);
        pw.cpix = cpix;
        pw.lmin = SD_FAR_LEVELS;
        pw.sat = 0;
        pw.pad = 0;
        // Far-capable pair: all parameters finite, a lowest level lmin at which it (and every level above) passes
        // level_ok, and a window of at least four level-lmin tiles (>= 512 pixels): only then can a tile two away from
        // the centre be covered.  Such pairs form class SD_FC0 + lmin and their window edges go to the edge list.
        if (fg.enabled && (hi - lo >= 512) && (r.thr == r.thr) && (a == a) && (fabs(a) < 1e300)) {
            const int n_act = fg.lev_info[0];
            int lmin = SD_FAR_LEVELS;
            const double gl = y * dw;
            for (int k = n_act - 1; k >= 0; k--) {
                if (!level_ok(fg, k, N, nus, cpix, r.nu, dw, r.inv_dw, gl, r.thr)) break;
                lmin = k;
            }
            if (lmin < n_act && hi - lo >= 4LL * fg.tile[lmin]) {
                cls = SD_FC0 + lmin;
                pw.lmin = (unsigned char)lmin;
                unsigned sat = 0;
                for (int k = lmin; k < n_act; k++) {
                    long long nb0 = 0, nb1 = N;
                    if (k + 1 < n_act) {  // the parent tile and its two neighbours
                        const long long P = cpix >> fg.tile_shift[k + 1], Tp = fg.tile[k + 1];
                        nb0 = (P - 1) * Tp > 0 ? (P - 1) * Tp : 0;
                        nb1 = (P + 2) * Tp < N ? (P + 2) * Tp : N;
                    }
                    if (lo <= nb0 && hi >= nb1) sat |= 1u << k;
                }
                pw.sat = (unsigned char)sat;
                // window edges strictly inside the extended range (tiles only look for edges strictly inside themselves)
                e_lo = lo > fg.ext0 && lo < fg.ext1;
                e_hi = hi < N && hi > fg.ext0 && hi < fg.ext1;
                key_lo = sd_edge_key(fg, 0, d, lmin, lo, (int)l);
                key_hi = sd_edge_key(fg, 1, d, lmin, hi, (int)l);
            }
        }
        s_wcls[td][tl] = (uint8_t)cls;
        pw.cls = (unsigned char)cls;
        s_win[td][tl] = pw;
        nonempty = hi > lo;
        wide = (hi > lo) && cls > 0;
        zero_dw = (hi > lo) && (dw == 0.0);
    }
    // ---- block-level aggregation: a handful of global counters are shared by all 65 000 blocks of this kernel, and one
    // atomic per WARP on each of them serialised in L2 (ncu: the kernel was waiting there, not on memory or math)
    __shared__ unsigned s_cnt[3];
    __shared__ unsigned s_edges[256 / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned m_lo = __ballot_sync(0xffffffffu, e_lo), m_hi = __ballot_sync(0xffffffffu, e_hi);
    const int n_lo = __popc(m_lo), n_hi = __popc(m_hi);
    if (fg.enabled && lane == 0) s_edges[wid] = (unsigned)(n_lo + n_hi);
    const unsigned ne_w = __popc(__ballot_sync(0xffffffffu, nonempty));
    const unsigned wd_w = __popc(__ballot_sync(0xffffffffu, wide));
    const unsigned zd_w = __popc(__ballot_sync(0xffffffffu, zero_dw));
    if (lane == 0) {
        if (ne_w) atomicAdd(&s_cnt[0], ne_w);
        if (wd_w) atomicAdd(&s_cnt[1], wd_w);
        if (zd_w) atomicAdd(&s_cnt[2], zd_w);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&stats[4], (unsigned long long)s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&stats[5], (unsigned long long)s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&stats[6], (unsigned long long)s_cnt[2]);
        if (fg.enabled) {
            unsigned tot = 0;
            for (int w2 = 0; w2 < 256 / 32; w2++) { const unsigned c = s_edges[w2]; s_edges[w2] = tot; tot += c; }
            s_base = tot ? atomicAdd(fg.edge_count, (unsigned long long)tot) : 0ull;
        }
    }
    __syncthreads();
    if (fg.enabled) {
        // append of the edge keys (order irrelevant: the keys are sorted into a total order)
        const unsigned long long base = s_base + s_edges[wid];
        if (e_lo) fg.edge_out[base + __popc(m_lo & lt)] = key_lo;
        if (e_hi) fg.edge_out[base + n_lo + __popc(m_hi & lt)] = key_hi;
    }
    // ---- the tile's rows (the barriers above ordered the shared-memory writes): 16-byte chunks, consecutive threads on
    // consecutive chunks of a depth row
    const int nl = (int)min((unsigned)TL, (unsigned)L - l0), nd = min(TD, D - d0);
    for (int c = threadIdx.x; c < TD * TL * 4; c += blockDim.x) {        // records: 4 chunks each
        const int row = c / (TL * 4), col = c % (TL * 4);
        if (row < nd && col < nl * 4)
            reinterpret_cast<int4 *>(rec + (size_t)(d0 + row) * L + l0)[col] = reinterpret_cast<const int4 *>(&s_rec[row][0])[col];
    }
    {
        const int row = threadIdx.x / TL, col = threadIdx.x % TL;            // window records: 1 chunk each; classes: 1 byte
        if (row < nd && col < nl) {
            reinterpret_cast<int4 *>(win + (size_t)(d0 + row) * L + l0)[col] = reinterpret_cast<const int4 *>(&s_win[row][0])[col];
            win_cls[(size_t)(d0 + row) * L + l0 + col] = s_wcls[row][col];
        }
    }
}

// ---- stable per-depth partition of the lines of class >= 1 by class (three small kernels) ----------
constexpr int CHUNK = 256;

__global__ void __launch_bounds__(CHUNK) k_cls_count(int64_t L, const uint8_t *__restrict__ win_cls, int nchunks,
                                                     int *__restrict__ chunk_cnt) {
    __shared__ int cnt[SD_NCLS];
    int d = blockIdx.y, ch = blockIdx.x;
    if (threadIdx.x < SD_NCLS) cnt[threadIdx.x] = 0;
    __syncthreads();
    int64_t l = (int64_t)ch * CHUNK + threadIdx.x;
    int cls = (l < L) ? win_cls[(size_t)d * L + l] : 0;
    for (int k = 1; k < SD_NCLS; k++) {
        unsigned b = __ballot_sync(0xffffffffu, cls == k);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt[k], __popc(b));
    }
    __syncthreads();
    if (threadIdx.x < SD_NCLS) chunk_cnt[((size_t)d * SD_NCLS + threadIdx.x) * nchunks + ch] = cnt[threadIdx.x];
}

// exclusive scan over (class, chunk) for each depth; one block per depth.
__global__ void __launch_bounds__(256) k_cls_scan(int nchunks, int *__restrict__ chunk_cnt, int *__restrict__ cls_off) {
    __shared__ int warp_tot[8];
    __shared__ int carry;
    int d = blockIdx.x;
    int *row = chunk_cnt + (size_t)d * SD_NCLS * nchunks;
    int n = SD_NCLS * nchunks;  // class-major order == order in cls_list
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        int i = base + threadIdx.x;
        int v = (i < n) ? row[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        int off = carry;
        for (int w = 0; w < (threadIdx.x >> 5); w++) off += warp_tot[w];
        int excl = off + x - v;
        if (i < n) {
            row[i] = excl;
            if (i % nchunks == 0) cls_off[d * (SD_NCLS + 1) + i / nchunks] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 255) carry = off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) cls_off[d * (SD_NCLS + 1) + SD_NCLS] = carry;
}

__global__ void __launch_bounds__(CHUNK) k_cls_scatter(int64_t L, const uint8_t *__restrict__ win_cls, int nchunks,
                                                       const int *__restrict__ chunk_off, int *__restrict__ cls_list) {
    __shared__ int warp_cnt[SD_NCLS][CHUNK / 32];
    int d = blockIdx.y, ch = blockIdx.x;
    int64_t l = (int64_t)ch * CHUNK + threadIdx.x;
    int cls = (l < L) ? win_cls[(size_t)d * L + l] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int my_rank = 0;
    for (int k = 1; k < SD_NCLS; k++) {
        unsigned b = __ballot_sync(0xffffffffu, cls == k);
        if (lane == 0) warp_cnt[k][w] = __popc(b);
        if (cls == k) my_rank = __popc(b & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (cls > 0) {
        int off = chunk_off[((size_t)d * SD_NCLS + cls) * nchunks + ch];
        for (int ww = 0; ww < w; ww++) off += warp_cnt[cls][ww];
        cls_list[(size_t)d * L + off + my_rank] = (int)l;
    }
}

}  // namespace

int sd_k1_broadening(sd_ctx *c, uint32_t flags) {
    int64_t n = c->L * c->D;
    SD_TRY(sd_ensure(c, c->gammas, sizeof(double) * n));
    SD_TRY(sd_ensure(c, c->dws, sizeof(double) * n));
    if (n == 0) return SD_OK;
    const double *stark = c->has_vald_cols ? c->l_stark.as<double>() : nullptr;
    const double *waals = c->has_vald_cols ? c->l_waals.as<double>() : nullptr;
    SD_TRY(sd_ensure(c, c->line_pre, sizeof(double) * LP_N * c->L));
    SD_TRY(sd_ensure(c, c->depth_pre, sizeof(double) * DP_N * c->D));
    k_line_pre<<<(unsigned)((c->L + 255) / 256), 256, 0, c->stream>>>(
        c->L, flags, c->l_nu.as<double>(), c->l_Z.as<int64_t>(), c->l_ion.as<int64_t>(), c->l_eion.as<double>(),
        c->l_eup.as<double>(), c->l_elo.as<double>(), c->l_A.as<double>(), c->l_mass.as<double>(), stark, waals,
        c->line_pre.as<double>());
    SD_TRY(sd_launch_check(c, "k_line_pre"));
    k_depth_pre<<<(c->D + 63) / 64, 64, 0, c->stream>>>(c->D, c->T.as<double>(), c->ne.as<double>(), c->nH.as<double>(),
                                                       c->depth_pre.as<double>());
    SD_TRY(sd_launch_check(c, "k_depth_pre"));
    const size_t smem = sizeof(double) * DP_N * c->D;
    SD_CHECK(c, smem <= 48 * 1024, SD_ERR_ARG, "sd_calc_broadening: too many depth points (%d)", c->D);
    k_broadening<<<(unsigned)((n + 255) / 256), 256, smem, c->stream>>>(c->L, c->D, c->line_pre.as<double>(),
                                                                       c->depth_pre.as<double>(), c->vmic, flags,
                                                                       c->gammas.as<double>(), c->dws.as<double>());
    return sd_launch_check(c, "k_broadening");
}

int sd_k2_prepare(sd_ctx *c) {
    int64_t L = c->L, n = c->L * c->D;
    int D = c->D;
    SD_TRY(sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS));
    SD_CUDA(c, cudaMemsetAsync(c->stats.p, 0, sizeof(unsigned long long) * 11, c->stream));  // [11] belongs to the strength producers
    sd_phase_begin(c, SD_PH_PREP);
    SD_TRY(sd_ensure(c, c->d_nu, sizeof(double)));
    SD_TRY(sd_ensure(c, c->cls_off, sizeof(int) * D * (SD_NCLS + 1)));
    k_dnu<<<1, 1024, 0, c->stream>>>(c->N, c->nus.as<double>(), c->d_nu.as<double>());
    SD_TRY(sd_launch_check(c, "k_dnu"));
    // pixels per thread of the line kernel and the geometry of the tile hierarchy (64 * 8^k pixels, fixed: the tiles
    // decide the summation order inside a pixel and must not depend on the launch configuration or the shard)
    c->k2_P = sd_k2_choose_P(c);
    FarGeom &fg = c->far_geom;
    for (int k = 0; k < SD_FAR_LEVELS; k++) {
        fg.tile_shift[k] = SD_FAR_TILE0_SHIFT + SD_FAR_SHIFT * k;
        fg.tile[k] = 1 << fg.tile_shift[k];
        fg.n_tiles[k] = (int)((c->N + fg.tile[k] - 1) / fg.tile[k]);
        SD_TRY(sd_ensure(c, c->tile_geom[k], sizeof(double) * 3 * fg.n_tiles[k]));
        fg.geom[k] = c->tile_geom[k].as<double>();
        k_tile_geometry<<<(fg.n_tiles[k] + 255) / 256, 256, 0, c->stream>>>(c->N, fg.tile[k], fg.n_tiles[k], c->nus.as<double>(),
                                                                          c->tile_geom[k].as<double>());
        SD_TRY(sd_launch_check(c, "k_tile_geometry"));
    }
    fg.lev_info = nullptr;
    fg.enabled = 0;
    fg.edge_keys = nullptr;
    fg.edge_off = nullptr;
    fg.fc_tab = nullptr;
    fg.edge_tab = nullptr;
    fg.edge_out = nullptr;
    fg.edge_count = nullptr;
    fg.l_bits = fg.pix_bits = fg.depth_bits = 1;
    // extended pixel range of this context: the top-level tiles its range [p0, p1) touches
    const long long T_top = fg.tile[SD_FAR_LEVELS - 1];
    fg.ext0 = (c->p0 / T_top) * T_top;
    fg.ext1 = ((c->p1 + T_top - 1) / T_top) * T_top;
    if (fg.ext1 > c->N) fg.ext1 = c->N;
    if (L == 0) {
        SD_CUDA(c, cudaMemsetAsync(c->cls_off.p, 0, sizeof(int) * D * (SD_NCLS + 1), c->stream));
        sd_phase_end(c, SD_PH_PREP);
        c->records_ready = true;
        c->far_active = 0;
        return SD_OK;
    }
    SD_TRY(sd_ensure(c, c->line_idx, sizeof(int) * L));
    SD_TRY(sd_ensure(c, c->rec, sizeof(LineRec) * n));
    SD_TRY(sd_ensure(c, c->win, sizeof(PairWin) * n));
    SD_TRY(sd_ensure(c, c->win_cls, n));
    SD_TRY(sd_ensure(c, c->cls_list, sizeof(int) * n));
    if (c->farfield) {
        fg.enabled = 1;
        while ((1LL << fg.pix_bits) <= c->N) fg.pix_bits++;
        while ((1 << fg.depth_bits) < D) fg.depth_bits++;
        while ((1LL << fg.l_bits) < L) fg.l_bits++;
        SD_CHECK(c, 1 + fg.depth_bits + SD_FAR_LMIN_BITS + fg.pix_bits + fg.l_bits <= 64, SD_ERR_ARG,
                 "far-field scheme: (depth, pixel, line) does not fit a 64-bit sort key (D = %d, N = %lld, L = %lld); use "
                 "sd_set_farfield(ctx, 0)", D, (long long)c->N, (long long)L);
        SD_TRY(sd_ensure(c, c->lev_info, sizeof(int) * (SD_FAR_LEVELS + 1)));
        k_level_check<<<1, 256, 0, c->stream>>>(fg, c->N, c->lev_info.as<int>());
        SD_TRY(sd_launch_check(c, "k_level_check"));
        fg.lev_info = c->lev_info.as<int>();
        SD_TRY(sd_ensure(c, c->edge_unsorted, sizeof(unsigned long long) * 2 * n));  // worst case: both edges of every pair
        SD_TRY(sd_ensure(c, c->edge_count, sizeof(unsigned long long)));
        SD_CUDA(c, cudaMemsetAsync(c->edge_count.p, 0, sizeof(unsigned long long), c->stream));
        fg.edge_out = c->edge_unsorted.as<unsigned long long>();
        fg.edge_count = c->edge_count.as<unsigned long long>();
    }
    int nchunks = (int)((L + CHUNK - 1) / CHUNK);
    SD_TRY(sd_ensure(c, c->chunk_cnt, sizeof(int) * (size_t)D * SD_NCLS * nchunks));
    k_line_idx<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(L, c->N, c->nus.as<double>(), c->l_nu.as<double>(),
                                                                  c->line_idx.as<int>());
    SD_TRY(sd_launch_check(c, "k_line_idx"));
    k_build_records<<<(unsigned)(((L + 31) / 32) * ((D + 7) / 8)), 256, 0, c->stream>>>(
        L, D, c->N, c->nus.as<double>(), c->l_nu.as<double>(), c->line_idx.as<int>(), c->gammas.as<double>(), c->gamma_cols,
        c->dws.as<double>(), c->l_alpha.as<double>(), c->d_nu.as<double>(), c->rec.as<LineRec>(), c->win.as<PairWin>(),
        c->win_cls.as<uint8_t>(), fg, c->stats.as<unsigned long long>());
    SD_TRY(sd_launch_check(c, "k_build_records"));
    k_cls_count<<<dim3(nchunks, D), CHUNK, 0, c->stream>>>(L, c->win_cls.as<uint8_t>(), nchunks, c->chunk_cnt.as<int>());
    SD_TRY(sd_launch_check(c, "k_cls_count"));
    k_cls_scan<<<D, 256, 0, c->stream>>>(nchunks, c->chunk_cnt.as<int>(), c->cls_off.as<int>());
    SD_TRY(sd_launch_check(c, "k_cls_scan"));
    k_cls_scatter<<<dim3(nchunks, D), CHUNK, 0, c->stream>>>(L, c->win_cls.as<uint8_t>(), nchunks, c->chunk_cnt.as<int>(),
                                                           c->cls_list.as<int>());
    SD_TRY(sd_launch_check(c, "k_cls_scatter"));
    sd_phase_end(c, SD_PH_PREP);
    if (c->farfield) {
        sd_phase_begin(c, SD_PH_SORT);
        SD_TRY(sd_sort_edges(c));
        sd_phase_end(c, SD_PH_SORT);
    }
    c->records_ready = true;
    return SD_OK;
}
