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

// Centre frequency and half-width [Hz] of every GLOBAL pixel tile [t*T, min((t+1)*T, N)) (far-field expansion point).
__global__ void __launch_bounds__(256) k_tile_geometry(int64_t N, int tile, int n_tiles, const double *__restrict__ nus,
                                                       double *__restrict__ geom) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int64_t a = (int64_t)t * tile, b = a + tile < N ? a + tile : N;
    double hi = nus[a], lo = nus[b - 1];
    geom[2 * t] = 0.5 * (hi + lo);
    geom[2 * t + 1] = 0.5 * (hi - lo);
}

// A tile is FAR from a (line, depth) pair when the Taylor expansion of the pair's region-I profile about the tile
// centre converges with ratio <= 1/SD_FAR_RHO_INV over the whole tile and every pixel of the tile is certainly in
// Humlicek region I.  Written as a positive condition: NaN parameters are never far.
__device__ __forceinline__ bool tile_is_far(double nu_c, double h, double nu_l, double dw, double y) {
    const double a_dw = 0.7071067811865476 * dw;  // the two poles of z/(z^2 - 1/2) sit at nu_l -+ dw/sqrt(2)
    double dist = fabs(nu_c - nu_l);
    double m = 15.0000001 - y;
    double core = h + (m > 0.0 ? m : 0.0) * dw;
    return (dist >= SD_FAR_RHO_INV * h + a_dw) && (dist >= core) && (dw > 0.0) && (y >= 0.0) && (y < 1e300) && (h > 0.0);
}

__device__ __forceinline__ int hw_class(long long hw) {
    // class 0: hw <= 64; class k (1..5): hw <= 64 * 4^k; class 6: everything wider (walked whole by the line kernel).
    // Class 7 (SD_FC_CLASS) is reserved for the far-capable pairs of the far-field scheme.
    if (hw <= SD_CLS0_HW) return 0;
    int k = 1;
    long long lim = (long long)SD_CLS0_HW * 4;
    while (k < SD_NCLS - 2 && hw > lim) { k++; lim *= 4; }
    return k;
}

// Tiles [a_, b_) around the centre tile tc that are NOT far.  Farness is monotone in the distance from the centre (the
// centre distance grows by 2 h per tile, the required distance by at most ~0.5 h), so each boundary is found from an
// estimate (required distance / tile spacing at the centre tile) plus a short walk.  The three tiles around either
// estimate are probed up front with INDEPENDENT loads (the walk then usually needs no further memory access: the
// dependent load -> test -> load chain of a plain walk is what kept the first version of this kernel latency bound).
__device__ __forceinline__ void near_interval(const double *__restrict__ geom, int n_tiles, int tc, double nu_l, double dw,
                                              double y, int &a_out, int &b_out) {
    const double h_c = geom[2 * tc + 1];
    const double m_c = 15.0000001 - y;
    const double need = fmax(SD_FAR_RHO_INV * h_c + 0.7071067811865476 * dw, h_c + (m_c > 0.0 ? m_c : 0.0) * dw);
    const double est = need / (2.0 * h_c);
    const int n_est = (est < (double)n_tiles) ? (int)est : n_tiles;  // NaN / inf -> the whole grid
    int a_ = max(tc - n_est, 0), b_ = min(tc + n_est + 1, n_tiles);
    // probes: tiles a_-1, a_, a_+1 and b_-2, b_-1, b_ (clamped; out-of-range probes are never consulted)
    const int pa = a_ - 1, pb = b_ - 2;
    bool fa[3], fb[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int ta = min(max(pa + k, 0), n_tiles - 1), tb = min(max(pb + k, 0), n_tiles - 1);
        fa[k] = tile_is_far(geom[2 * ta], geom[2 * ta + 1], nu_l, dw, y);
        fb[k] = tile_is_far(geom[2 * tb], geom[2 * tb + 1], nu_l, dw, y);
    }
    auto far_a = [&](int t) {
        const int k = t - pa;
        return (k >= 0 && k < 3) ? (k == 0 ? fa[0] : (k == 1 ? fa[1] : fa[2])) : tile_is_far(geom[2 * t], geom[2 * t + 1], nu_l, dw, y);
    };
    auto far_b = [&](int t) {
        const int k = t - pb;
        return (k >= 0 && k < 3) ? (k == 0 ? fb[0] : (k == 1 ? fb[1] : fb[2])) : tile_is_far(geom[2 * t], geom[2 * t + 1], nu_l, dw, y);
    };
    while (a_ > 0 && !far_a(a_ - 1)) a_--;
    while (a_ < tc && far_a(a_)) a_++;
    while (b_ < n_tiles && !far_b(b_)) b_++;
    while (b_ > tc + 1 && far_b(b_ - 1)) b_--;
    a_out = a_;
    b_out = b_;
}

// One thread per (line, depth), d fastest (coalesced reads of the (L,D) inputs).  Writes depth-major
// records/windows (64-byte records are two full sectors, so the transposing write is not wasteful).
__global__ void __launch_bounds__(256) k_build_records(int64_t L, int D, int64_t N, const double *__restrict__ line_nu,
                                                       const int *__restrict__ line_idx, const double *__restrict__ gammas,
                                                       int gamma_cols, const double *__restrict__ dws,
                                                       const double *__restrict__ alpha, const double *__restrict__ d_nu_p,
                                                       LineRec *__restrict__ rec, PairWin *__restrict__ win,
                                                       uint8_t *__restrict__ win_cls,
                                                       FarGeom fg, unsigned long long *__restrict__ stats) {
    const unsigned g = blockIdx.x * blockDim.x + threadIdx.x;  // L * D < 2^31 (checked by sd_set_lines)
    bool active = g < (unsigned)(L * D);
    unsigned nonempty = 0, wide = 0, zero_dw = 0;
    int rad_k[SD_FAR_LEVELS];
#pragma unroll
    for (int k = 0; k < SD_FAR_LEVELS; k++) rad_k[k] = 0;
    bool e_lo = false, e_hi = false;
    unsigned long long key_lo = 0, key_hi = 0;
    if (active) {
        const unsigned l = g / (unsigned)D;
        const int d = (int)(g - l * (unsigned)D);
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
        size_t o = (size_t)d * L + l;
        // The record is read only after a window test passed for a tile of this context: pairs whose window is empty
        // or misses the extended pixel range never get that far (nu sharding: ~half of the record traffic per rank).
        if (hi > lo && hi > fg.ext0 && lo < fg.ext1) rec[o] = r;
        PairWin pw;
        pw.lo = (int)lo;
        pw.hi = (int)hi;
        pw.near[0] = pw.near[1] = pw.near[2] = 0xffff0000u;
        pw.pad0 = pw.pad1 = 0;
        // Far-capable pair: the window holds at least one level-0 tile and all parameters are finite.  Such pairs
        // form class 7; per hierarchy level they get the interval of tiles [nl, nh) around the line centre that are
        // NOT far, and their window edges go to the edge list.
        const bool fc = fg.enabled && (hi - lo >= fg.tile[0]) && (r.thr == r.thr) && (a == a) && (fabs(a) < 1e300);
        if (fc) cls = SD_FC_CLASS;
        win_cls[o] = (uint8_t)cls;
        pw.cls = cls;
        if (fg.enabled) {
#pragma unroll
            for (int k = 0; k < SD_FAR_LEVELS; k++) {
                unsigned nl = 0, nh = 0xffffu;
                int rad = 0;
                const int tile = fg.tile[k], n_tiles = fg.n_tiles[k];
                if (fc && hi - lo >= tile) {
                    int tc = idx >> fg.tile_shift[k];  // tiles hold 2^tile_shift pixels
                    if (tc >= n_tiles) tc = n_tiles - 1;
                    int a_, b_;
                    near_interval(fg.geom[k], n_tiles, tc, r.nu, dw, y, a_, b_);
                    nl = (unsigned)a_;
                    nh = (unsigned)b_;
                    rad = max(tc - a_, b_ - 1 - tc);
                }
                pw.near[k] = nl | (nh << 16);
                rad_k[k] = rad;
            }
            // window edges strictly inside the extended range (tiles only look for edges strictly inside themselves)
            e_lo = fc && lo > fg.ext0 && lo < fg.ext1;
            e_hi = fc && hi < N && hi > fg.ext0 && hi < fg.ext1;
            key_lo = sd_edge_key(fg, 0, d, lo, (int)l);
            key_hi = sd_edge_key(fg, 1, d, hi, (int)l);
        }
        win[o] = pw;
        nonempty = hi > lo;
        wide = (hi > lo) && cls > 0;
        zero_dw = (hi > lo) && (dw == 0.0);
    }
    // ---- block-level aggregation: a handful of global counters are shared by all 65 000 blocks of this kernel, and one
    // atomic per WARP on each of them serialised in L2 (ncu: the kernel was waiting there, not on memory or math)
    __shared__ int s_rad[SD_FAR_LEVELS];
    __shared__ unsigned s_cnt[3];
    __shared__ unsigned s_edges[256 / 32];
    __shared__ unsigned long long s_base;
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    if (threadIdx.x < SD_FAR_LEVELS) s_rad[threadIdx.x] = 0;
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const unsigned m_lo = __ballot_sync(0xffffffffu, e_lo), m_hi = __ballot_sync(0xffffffffu, e_hi);
    const int n_lo = __popc(m_lo), n_hi = __popc(m_hi);
    if (fg.enabled) {
#pragma unroll
        for (int k = 0; k < SD_FAR_LEVELS; k++) {
            int m = rad_k[k];
            for (int o2 = 16; o2; o2 >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o2));
            if (lane == 0 && m > 0) atomicMax(&s_rad[k], m);
        }
        if (lane == 0) s_edges[wid] = (unsigned)(n_lo + n_hi);
    }
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
            // the running maximum is reached early: most blocks find nothing to raise (racy read, monotone update)
            for (int k = 0; k < SD_FAR_LEVELS; k++)
                if (s_rad[k] > *(volatile int *)&fg.near_rad[k]) atomicMax(&fg.near_rad[k], s_rad[k]);
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
    // pixels per thread of the line kernel (level-0 tile = 256 * P pixels) and the geometry of the tile hierarchy
    c->k2_P = sd_k2_choose_P(c);
    FarGeom &fg = c->far_geom;
    for (int k = 0; k < SD_FAR_LEVELS; k++) {
        fg.tile[k] = (32 * c->k2_NW * c->k2_P) << (SD_FAR_SHIFT * k);
        fg.tile_shift[k] = 0;
        while ((1 << fg.tile_shift[k]) < fg.tile[k]) fg.tile_shift[k]++;
        SD_CHECK(c, (1 << fg.tile_shift[k]) == fg.tile[k], SD_ERR_STATE, "tile sizes must be powers of two");
        fg.n_tiles[k] = (int)((c->N + fg.tile[k] - 1) / fg.tile[k]);
        SD_CHECK(c, fg.n_tiles[k] < 65535, SD_ERR_ARG, "grid too long for 16-bit tile indices");
        SD_TRY(sd_ensure(c, c->tile_geom[k], sizeof(double) * 2 * fg.n_tiles[k]));
        fg.geom[k] = c->tile_geom[k].as<double>();
        k_tile_geometry<<<(fg.n_tiles[k] + 255) / 256, 256, 0, c->stream>>>(c->N, fg.tile[k], fg.n_tiles[k], c->nus.as<double>(),
                                                                          c->tile_geom[k].as<double>());
        SD_TRY(sd_launch_check(c, "k_tile_geometry"));
    }
    fg.near_rad = nullptr;
    fg.enabled = 0;
    fg.edge_keys = nullptr;
    fg.edge_off = nullptr;
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
        SD_CHECK(c, 1 + fg.depth_bits + fg.pix_bits + fg.l_bits <= 64, SD_ERR_ARG,
                 "far-field scheme: (depth, pixel, line) does not fit a 64-bit sort key (D = %d, N = %lld, L = %lld); use "
                 "sd_set_farfield(ctx, 0)", D, (long long)c->N, (long long)L);
        SD_TRY(sd_ensure(c, c->near_rad, sizeof(int) * SD_FAR_LEVELS));
        SD_CUDA(c, cudaMemsetAsync(c->near_rad.p, 0, sizeof(int) * SD_FAR_LEVELS, c->stream));
        fg.near_rad = c->near_rad.as<int>();
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
    k_build_records<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        L, D, c->N, c->l_nu.as<double>(), c->line_idx.as<int>(), c->gammas.as<double>(), c->gamma_cols,
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
