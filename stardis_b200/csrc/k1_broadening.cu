// k1_broadening.cu -- K1: collisional broadening and Doppler widths per (line, depth), and the K2
// preparation pass (global window of every (line, depth) pair, half-width classes, 64-byte line records).
//
// Reference: stardis/radiation_field/opacities/opacities_solvers/broadening.py:550-656 (calc_gamma),
// :659-732 (calculate_broadening: z_eff = ion_number + 1, linear Stark for hydrogen only), :1009-1085
// (calc_vald_gamma), :32-71 (calc_doppler_width); opacities_solvers/base.py:522-575 (window rule).
//
// Roofline: HBM.  Per (line, depth): reads alpha_line 8 B (+ O(L) per-line columns), writes gamma 8 B,
// doppler width 8 B, record 64 B, window 9 B.
#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

__global__ void __launch_bounds__(256) k_broadening(int64_t L, int D, const double *__restrict__ nu,
                                                    const int64_t *__restrict__ Z, const int64_t *__restrict__ ion,
                                                    const double *__restrict__ e_ion, const double *__restrict__ e_up,
                                                    const double *__restrict__ e_lo, const double *__restrict__ A_ul,
                                                    const double *__restrict__ mass, const double *__restrict__ stark,
                                                    const double *__restrict__ waals, const double *__restrict__ T,
                                                    const double *__restrict__ ne, const double *__restrict__ nH, double vmic,
                                                    uint32_t flags, double *__restrict__ gammas, double *__restrict__ dws) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // (l, d), d fastest
    if (g >= L * D) return;
    int64_t l = g / D;
    int d = (int)(g - l * D);
    double zeff = (double)(ion[l] + 1);
    double n_up = sdm::n_effective(zeff, e_ion[l], e_up[l]);
    double n_lo = sdm::n_effective(zeff, e_ion[l], e_lo[l]);
    double Td = T[d], ned = ne[d], nHd = nH[d];
    double gam;
    if (flags & SD_VALD) {  // broadening.py:1039-1085
        gam = 0.0;
        if (flags & SD_RADIATION) gam += A_ul[l];
        if ((flags & SD_LINEAR_STARK) && Z[l] == 1) gam += sdm::gamma_linear_stark(n_up, n_lo, ned);
        if (flags & SD_QUADRATIC_STARK) gam += sdm::vald_stark(ned, stark[l], Td);
        if (flags & SD_VAN_DER_WAALS) gam += sdm::vald_vdw_unit(waals[l], zeff, n_up, n_lo, Td, mass[l]) * nHd;
        gam /= 2.0;
    } else {  // broadening.py:611-654
        double g_ls = 0.0, g_qs = 0.0, g_vdw = 0.0, g_rad = 0.0;
        if ((flags & SD_LINEAR_STARK) && Z[l] == 1) g_ls = sdm::gamma_linear_stark(n_up, n_lo, ned);
        if (flags & SD_QUADRATIC_STARK) g_qs = sdm::gamma_quadratic_stark(zeff, n_up, n_lo, ned, Td);
        if (flags & SD_VAN_DER_WAALS) g_vdw = sdm::gamma_van_der_waals(zeff, n_up, n_lo, Td, nHd);
        if (flags & SD_RADIATION) g_rad = A_ul[l];
        gam = g_ls + g_qs + g_vdw + g_rad;
    }
    gammas[g] = gam;
    dws[g] = sdm::doppler_width(nu[l], Td, mass[l], vmic);
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

// One thread per (line, depth), d fastest (coalesced reads of the (L,D) inputs).  Writes depth-major
// records/windows (64-byte records are two full sectors, so the transposing write is not wasteful).
__global__ void __launch_bounds__(256) k_build_records(int64_t L, int D, int64_t N, const double *__restrict__ line_nu,
                                                       const int *__restrict__ line_idx, const double *__restrict__ gammas,
                                                       int gamma_cols, const double *__restrict__ dws,
                                                       const double *__restrict__ alpha, const double *__restrict__ d_nu_p,
                                                       LineRec *__restrict__ rec, PairWin *__restrict__ win,
                                                       uint8_t *__restrict__ win_cls,
                                                       FarGeom fg, unsigned long long *__restrict__ stats) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool active = g < L * D;
    unsigned nonempty = 0, wide = 0, zero_dw = 0;
    int rad_k[SD_FAR_LEVELS];
#pragma unroll
    for (int k = 0; k < SD_FAR_LEVELS; k++) rad_k[k] = 0;
    if (active) {
        int64_t l = g / D;
        int d = (int)(g - l * D);
        double gam = gamma_cols > 1 ? gammas[g] : gammas[l];  // base.py:547-551
        double dw = dws[g];
        double a = alpha[g];
        double d_nu = d_nu_p[0];
        long long lo, hi;
        sdm::line_window(line_idx[l], N, gam, dw, a, d_nu, lo, hi);
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
        rec[o] = r;
        PairWin pw;
        pw.lo = (int)lo;
        pw.hi = (int)hi;
        pw.near[0] = pw.near[1] = pw.near[2] = 0xffff0000u;
        pw.pad0 = pw.pad1 = 0;
        // Far-capable pair: the window holds at least one level-0 tile and all parameters are finite.  Such pairs
        // form class 7; per hierarchy level they get the interval of tiles [nl, nh) around the line centre that are
        // NOT far, and their window edges go to the two edge-sort key arrays.
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
                    const double *__restrict__ geom = fg.geom[k];
                    int tc = (int)(line_idx[l] / tile);
                    if (tc >= n_tiles) tc = n_tiles - 1;
                    // [a_, b_): tiles around the line centre that are not far.  Farness is monotone in the distance from
                    // the centre (the centre distance grows by 2 h per tile, the required distance by at most ~0.5 h), so
                    // the two boundaries are found from an estimate (required distance / tile spacing at the centre
                    // tile) plus a short walk in either direction instead of a walk from the centre.
                    const double h_c = geom[2 * tc + 1];
                    const double m_c = 15.0000001 - y;
                    const double need = fmax(SD_FAR_RHO_INV * h_c + 0.7071067811865476 * dw, h_c + (m_c > 0.0 ? m_c : 0.0) * dw);
                    const double est = need / (2.0 * h_c);
                    const int n_est = (est < (double)n_tiles) ? (int)est : n_tiles;  // NaN / inf -> the whole grid
                    int a_ = max(tc - n_est, 0), b_ = min(tc + n_est + 1, n_tiles);
                    while (a_ > 0 && !tile_is_far(geom[2 * (a_ - 1)], geom[2 * (a_ - 1) + 1], r.nu, dw, y)) a_--;
                    while (a_ < tc && tile_is_far(geom[2 * a_], geom[2 * a_ + 1], r.nu, dw, y)) a_++;
                    while (b_ < n_tiles && !tile_is_far(geom[2 * b_], geom[2 * b_ + 1], r.nu, dw, y)) b_++;
                    while (b_ > tc + 1 && tile_is_far(geom[2 * (b_ - 1)], geom[2 * (b_ - 1) + 1], r.nu, dw, y)) b_--;
                    nl = (unsigned)a_;
                    nh = (unsigned)b_;
                    rad = max(tc - a_, b_ - 1 - tc);
                }
                pw.near[k] = nl | (nh << 16);
                rad_k[k] = rad;
            }
            const unsigned dkey = (unsigned)d << fg.key_shift, none = (1u << fg.key_shift) - 1u;
            fg.lo_keys[o] = dkey | ((fc && lo > 0) ? (unsigned)lo : none);
            fg.hi_keys[o] = dkey | ((fc && hi < N) ? (unsigned)hi : none);
            fg.lo_l[o] = (int)l;
        }
        win[o] = pw;
        nonempty = hi > lo;
        wide = (hi > lo) && cls > 0;
        zero_dw = (hi > lo) && (dw == 0.0);
    }
    if (fg.enabled) {
#pragma unroll
        for (int k = 0; k < SD_FAR_LEVELS; k++) {
            int m = rad_k[k];
            for (int o2 = 16; o2; o2 >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o2));
            if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&fg.near_rad[k], m);
        }
    }
    unsigned ne_w = __popc(__ballot_sync(0xffffffffu, nonempty));
    unsigned wd_w = __popc(__ballot_sync(0xffffffffu, wide));
    unsigned zd_w = __popc(__ballot_sync(0xffffffffu, zero_dw));
    if ((threadIdx.x & 31) == 0) {
        if (ne_w) atomicAdd(&stats[4], (unsigned long long)ne_w);
        if (wd_w) atomicAdd(&stats[5], (unsigned long long)wd_w);
        if (zd_w) atomicAdd(&stats[6], (unsigned long long)zd_w);
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
    k_broadening<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->L, c->D, c->l_nu.as<double>(), c->l_Z.as<int64_t>(), c->l_ion.as<int64_t>(), c->l_eion.as<double>(),
        c->l_eup.as<double>(), c->l_elo.as<double>(), c->l_A.as<double>(), c->l_mass.as<double>(), stark, waals,
        c->T.as<double>(), c->ne.as<double>(), c->nH.as<double>(), c->vmic, flags, c->gammas.as<double>(),
        c->dws.as<double>());
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
    fg.key_shift = 0;
    fg.lo_keys = fg.hi_keys = nullptr;
    fg.lo_l = fg.hi_l = nullptr;
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
        int pix_bits = 1, depth_bits = 1;
        while ((1LL << pix_bits) <= c->N) pix_bits++;
        while ((1 << depth_bits) < D) depth_bits++;
        SD_CHECK(c, pix_bits + depth_bits <= 32, SD_ERR_ARG,
                 "far-field scheme: (depth, pixel) does not fit a 32-bit sort key (D = %d, N = %lld); use sd_set_farfield(ctx, 0)",
                 D, (long long)c->N);
        fg.key_shift = pix_bits;
        SD_TRY(sd_ensure(c, c->near_rad, sizeof(int) * SD_FAR_LEVELS));
        SD_CUDA(c, cudaMemsetAsync(c->near_rad.p, 0, sizeof(int) * SD_FAR_LEVELS, c->stream));
        fg.near_rad = c->near_rad.as<int>();
        // unsorted keys go to the temporaries (window starts) / to edge_keys[1] (window ends, sorted second)
        SD_TRY(sd_ensure(c, c->edge_tmp_keys, sizeof(unsigned) * n));
        SD_TRY(sd_ensure(c, c->edge_tmp_l, sizeof(int) * n));
        for (int w = 0; w < 2; w++) {
            SD_TRY(sd_ensure(c, c->edge_keys[w], sizeof(unsigned) * n));
            SD_TRY(sd_ensure(c, c->edge_l[w], sizeof(int) * n));
        }
        fg.lo_keys = c->edge_tmp_keys.as<unsigned>();
        fg.hi_keys = c->edge_keys[0].as<unsigned>();  // staging; overwritten by the first sort's output later
        fg.lo_l = c->edge_tmp_l.as<int>();
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
        // window ends were staged in edge_keys[0]: sort them first (into edge_keys[1]/edge_l[1]), then the starts
        SD_TRY(sd_sort_edges(c, 1, n));
        SD_TRY(sd_sort_edges(c, 0, n));
        fg.lo_keys = c->edge_keys[0].as<unsigned>();
        fg.lo_l = c->edge_l[0].as<int>();
        fg.hi_keys = c->edge_keys[1].as<unsigned>();
        fg.hi_l = c->edge_l[1].as<int>();
        sd_phase_end(c, SD_PH_SORT);
    }
    c->records_ready = true;
    return SD_OK;
}
