// k2_lines.cu -- K2: windowed Voigt accumulation  alpha_line[d, i] = sum_l phi(nu_i - nu_l; dw[l,d], gamma[l,d]) * alpha[l,d]
// over the pixels i in the (line, depth) window [lo, hi) only.
//
// Reference: stardis/radiation_field/opacities/opacities_solvers/base.py:487-627 (calc_alan_entries, thread-
// parallel over lines with one private (D,N) slab per numba thread) and voigt.py:17-155.
//
// B200 design (gather, no atomics, deterministic):
//   * one CTA owns a GLOBAL tile of 256*P consecutive pixels of ONE depth point (tiles are aligned to the global grid
//     so that a nu shard reproduces the full-grid result bit for bit) and keeps the P accumulators of every thread in
//     registers for the whole kernel; the result is written exactly once, coalesced;
//   * the (line, depth) pairs that can touch the tile are found per half-width class (class 0: contiguous range of
//     the nu-sorted line list; class k >= 1: contiguous range of the per-depth class list built by k1_broadening.cu;
//     32-ary warp binary searches on the monotone window centres) and, for the far-capable classes, from range
//     tables per 64-pixel tile boundary (two loads);
//   * k_lines: every WARP streams the candidates in batches of 32 on its own (no CTA barrier in the main loop): test
//     against the warp's 32*P-pixel span, expand the passing (line, depth) records into 96-byte shared-memory entries
//     (constants hoisted once per warp), entries whose window covers the span and whose span lies entirely in
//     Humlicek region I packed first ("far-wing" list), the others from the back ("mixed" list), then consume them
//     with broadcast LDS; summation order per pixel is fixed;
//   * far-wing hot loop (region I):  Kf (q + c1) / (q (q + b) + c),  q = x^2:  8 FP64 instructions + 1 MUFU.RCP64H
//     per evaluation (x, q, 2 for the denominator, numerator, 2 for the Newton step on the reciprocal seed,
//     accumulate), no branches, no divisions;
//   * pixels that are not certainly in region I take the exact path: x = dnu / dw (correctly rounded) and the
//     reference's own region tests, so the Humlicek classification is identical to the reference's;
//   * FAR FIELD (k_s2m, k_m2m, k_m2l, k_far_coeffs + polynomial epilogue of k_lines): in region I,
//     Re w = (1/(2 sqrt(pi))) [ y/((x-a)^2+y^2) + y/((x+a)^2+y^2) ],  a = 1/sqrt(2): two Lorentzians, i.e. the
//     imaginary part of two simple poles p = nu_l -+ dw/sqrt(2) + i y dw with a real weight.  On a hierarchy of pixel
//     tiles (64 * 8^k pixels) a pair whose window covers a tile at least two tiles away from its centre contributes a
//     function that is analytic over that tile: the degree-31 Taylor polynomials of all such pairs of a tile are
//     summed once -- through multipole moments of the source tiles and real tile-to-tile translation matrices (a 1-D
//     fast multipole method) for pairs whose window covers the whole neighbourhood, by direct expansion otherwise --
//     k_lines skips exactly those (pair, 64-pixel tile) products (same integer test on the 16-byte window record)
//     and adds the polynomials at the end.  See sd_internal.h and the kernels below.
//
// Roofline: FP64 FMA pipe (no dense contraction -> no tensor cores).  Memory traffic is negligible.
#include <stdlib.h>

#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

int env_int_early(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int FAR_CH = 1024;                // candidates per cooperative scan chunk of k_far_coeffs
constexpr int K1 = SD_FAR_K + 1;            // coefficients per polynomial = multipole moments per tile
struct __align__(16) FarRec { double nu, dw, y, K; };  // = the first 32 bytes of LineRec
static_assert(K1 == 32, "k_m2l maps the lanes of a warp to the coefficients; k_s2m / k_far_coeffs reduce into lane k");
static_assert(K1 % 4 == 0, "the series length is tested every fourth term");
static_assert(WARPS == (1 << SD_FAR_SHIFT), "k_far_coeffs / k_s2m / k_m2l map the warps of a CTA to the children of a tile");
constexpr size_t FAR_SMEM = (size_t)FAR_CH * sizeof(FarRec) + FAR_CH;
constexpr int FAR_MAX_SRC = 3 * SD_FAR_LEVELS;  // candidate lists of one k_far_coeffs group
constexpr int M2L_DC_MAX = 14;              // depth points per k_m2l CTA (accumulators per thread); 8 when few depths are local

struct __align__(16) WEntry {
    // far-wing path (48 B)
    double xl;      // nu_l / dw
    double inv_dw;  // 1 / dw
    double b;       // 2 y^2 - 1
    double c;       // (y^2 + 1/2)^2
    double Kc;      // Kf (y^2 + 1/2)
    double Kf;      // alpha y / (pi dw)
    // window + region-I threshold (16 B)
    double thr;     // q > thr  =>  region I for certain
    int lo, hi;
    // exact path (32 B)
    double nu, dw, y, K;
};
static_assert(sizeof(WEntry) == 96, "WEntry layout");

struct LineArgs {
    int64_t L, N, p0, p1;
    int D;
    int tile0;     // global index of the first CTA tile of this launch
    int n_tiles;   // global number of CTA tiles
    const double *nus;
    const int *line_idx;
    const LineRec *rec;
    const PairWin *win;   // (depth, line) window records
    const int *cls_list, *cls_off;
    FarGeom fg;                  // tile hierarchy; fg.enabled == 0: far field disabled
    int n_act;                   // usable hierarchy levels (level n_act - 1 is the top level)
    double *far_coef[SD_FAR_LEVELS];   // per level: (D, n_tiles_launch[k], K1)
    int far_tile0[SD_FAR_LEVELS];      // first global tile of this launch, per level
    int far_ntl[SD_FAR_LEVELS];        // tiles of this launch, per level
    double *far_mom[SD_FAR_LEVELS];    // per level: (D, n_tiles[k], K1) multipole moments of the pairs saturated at level k
    // far_bkt[k][h], h >= k: (D, n_tiles[k], K1) moments about the level-k tile centres of the pairs with lmin <= k
    // whose HIGHEST saturated level is h (k_s2m fills level lmin, k_m2m translates upwards and sums far_mom)
    double *far_bkt[SD_FAR_LEVELS][SD_FAR_LEVELS];
    double *out;                 // (D, p1 - p0)
    unsigned long long *stats;
};

// smallest j in [a, b] with (j == b or key(j) < X); key non-increasing in j.  Warp-cooperative 32-ary search.
template <class KeyFn>
__device__ __forceinline__ int warp_first_below(KeyFn key, int a, int b, int X) {
    const int lane = threadIdx.x & 31;
    while (b > a) {
        int n = b - a;
        int step = (n + 31) >> 5;
        long long pj = (long long)a + (long long)lane * step;
        bool pred = (pj >= b) ? true : (key((int)pj) < X);
        unsigned m = __ballot_sync(0xffffffffu, pred);
        int f = m ? (__ffs(m) - 1) : 32;
        if (f == 0) return a;
        int na = a + (f - 1) * step + 1;
        long long nb = (f < 32) ? (long long)a + (long long)f * step : (long long)b;
        a = na;
        b = (int)(nb < b ? nb : b);
    }
    return a;
}

__device__ __forceinline__ int clamp_i32(long long v) {
    return (int)(v > 2147483647LL ? 2147483647LL : (v < -2147483647LL ? -2147483647LL : v));
}

// Candidate range [ja, jb) of half-width class `cls` for the pixel interval [t0, t1) at depth d; executed by one warp.
__device__ __forceinline__ void class_range(const LineArgs &a, int d, int cls, int64_t t0, int64_t t1, int &ja, int &jb) {
    const int *line_idx = a.line_idx;
    if (cls == 0) {
        auto key = [&](int j) { return line_idx[j]; };
        ja = warp_first_below(key, 0, (int)a.L, clamp_i32(t1 + SD_CLS0_HW));        // idx <  t1 + H
        jb = warp_first_below(key, ja, (int)a.L, clamp_i32(t0 - SD_CLS0_HW + 1));   // idx <= t0 - H
        return;
    }
    const int lo = a.cls_off[d * (SD_NCLS + 1) + cls], hi = a.cls_off[d * (SD_NCLS + 1) + cls + 1];
    if (cls >= SD_FC0 - 1) {  // class 6 (unbounded half-width): every pair is a candidate
        ja = lo;
        jb = hi;
        return;
    }
    const int *list_d = a.cls_list + (size_t)d * a.L;
    const long long H = (long long)SD_CLS0_HW << (2 * cls);
    auto key = [&](int j) { return line_idx[list_d[j]]; };
    ja = warp_first_below(key, lo, hi, clamp_i32(t1 + H));
    jb = warp_first_below(key, ja, hi, clamp_i32(t0 - H + 1));
}

// Far-capable pairs with lmin = m (class SD_FC0 + m, sorted by centre) whose centre pixel lies in [c0, c1), both
// multiples of the level-0 tile size (or beyond the grid): two loads from the range table (FarGeom::fc_tab).
__device__ __forceinline__ void fc_centre_range(const LineArgs &a, int d, int m, long long c0, long long c1, int &ja, int &jb) {
    const int nt0 = a.fg.n_tiles[0];
    const int *__restrict__ row = a.fg.fc_tab + (size_t)(d * SD_FAR_LEVELS + m) * (nt0 + 1);
    const long long ta = c0 <= 0 ? 0 : (c0 >> SD_FAR_TILE0_SHIFT), tb = c1 >= a.N ? nt0 : (c1 >> SD_FAR_TILE0_SHIFT);
    ja = row[tb < nt0 ? tb : nt0];
    jb = row[ta < nt0 ? ta : nt0];
}

// Far-capable pairs of depth d and lmin = m with a window start (which = 0) or end (which = 1) at a pixel in [t0, t1),
// t0 a multiple of the level-0 tile size, t1 one or the end of the grid: two loads from the range table
// (FarGeom::edge_tab).  Callers that need the edge STRICTLY inside (t0, t1) reject an edge at t0 themselves.
__device__ __forceinline__ void fc_edge_range(const LineArgs &a, int d, int which, int m, int64_t t0, int64_t t1, int &ja, int &jb) {
    const int nt0 = a.fg.n_tiles[0];
    const int *__restrict__ row = a.fg.edge_tab + (size_t)((which * a.D + d) * SD_FAR_LEVELS + m) * (nt0 + 1);
    const int64_t ta = t0 >> SD_FAR_TILE0_SHIFT, tb = t1 >= a.N ? nt0 : (t1 >> SD_FAR_TILE0_SHIFT);
    ja = row[ta < nt0 ? ta : nt0];
    jb = row[tb < nt0 ? tb : nt0];
}

// one 16-byte gather
__device__ __forceinline__ PairWin load_win(const PairWin *__restrict__ w) {
    const int4 a = __ldg(reinterpret_cast<const int4 *>(w));
    PairWin r;
    r.lo = a.x; r.hi = a.y; r.cpix = a.z;
    r.cls = (unsigned char)(a.w & 0xff); r.lmin = (unsigned char)((a.w >> 8) & 0xff); r.sat = (unsigned char)((a.w >> 16) & 0xff);
    r.pad = 0;
    return r;
}

// x = (nu_i - nu_l) / dw (voigt.py:148) must be the correctly rounded quotient: the W4 regions are chosen by comparing
// |x| + y with literals and the approximation jumps by ~1e-4 across a region boundary.  With the correctly rounded
// reciprocal r = RN(1 / dw) stored per pair, q = RN(n r) followed by one residual step q + (n - dw q) r is the correctly
// rounded quotient (Markstein) in 3 instead of ~25 instructions; pairs with dw <= 0, inf or NaN (thr is NaN for them)
// keep the IEEE division so that their special values propagate exactly as in the reference.
__device__ __noinline__ double exact_contribution(double nu_i, double nu_l, double dw, double inv_dw, double thr, double y,
                                                  double K) {
    const double n = nu_i - nu_l;
    double x;
    if (thr == thr) {
        const double q = n * inv_dw;
        x = fma(fma(-dw, q, n), inv_dw, q);
    } else {
        x = n / dw;
    }
    return sdm::humlicek_re(x, y) * K;
}

// The same integer test in all kernels: at level `lev` (tiles of 2^shift pixels) the pair is FAR from tile `t` =
// [t0, t1): expandable there, centre at least two tiles away, window covering the tile.  Farness by distance at one
// level implies it at every lower level >= lmin and covering a tile implies covering its children, so "far at some
// level" is decided at level lmin alone (k_lines) and "served at level lev" = far at lev and not far at lev + 1.
__device__ __forceinline__ bool pair_is_far(const PairWin &w, int lev, int shift, int t, int64_t t0, int64_t t1) {
    const int ds = (w.cpix >> shift) - t;
    return (lev >= (int)w.lmin) && (ds >= 2 || ds <= -2) && (w.lo <= t0) && (w.hi >= t1);
}

// ------------------------------------------------------------------------------------------------------------------
// Far-field coefficients of one (level-`lev` tile, depth), direct part: C_k = A Im(v w^k) summed over the two poles,
// v = 1 / (nu_c - p), w = -h v, A = K dw / (2 sqrt(pi));  the contribution of the pair at pixel nu is
// sum_k C_k ((nu - nu_c)/h)^k.  This kernel expands the pairs that are served at level `lev` (far there, not far for
// the parent tile) and are NOT saturated there -- the saturated ones go through k_s2m / k_m2l.  They are found without
// scanning, per lmin class m <= lev:
//   (A) pairs whose centre lies in the parent tile or one of its two neighbours (never far from the parent): a
//       contiguous range (by centre) of the class list; the saturated ones are skipped;
//   (B) pairs far from the parent by distance that do not cover it (a window edge lies strictly inside it): two ranges
//       of the edge-sorted lists.
// The candidates are a property of the PARENT, so one CTA serves the eight children of a parent (warp w = child w):
//   scan     the CTA walks the candidate lists in chunks of FAR_CH; a thread gathers the 16-byte window record of a
//            candidate ONCE, tests it against all eight children (integer compares) and leaves an 8-bit acceptance mask;
//            the 32 bytes of LineRec the expansion needs are staged in shared memory with cp.async if any child wants
//            them (one gather per candidate and parent instead of one per candidate and child);
//   expand   every warp walks the dense list 32 entries at a time, one pair per lane.
// The top level has no parent: groups of eight consecutive tiles walk a fixed slice of the whole class lists
// (`nsplit` CTAs per group; k_far_reduce adds the partial sums in slice order).  Groups, slices, chunking and list order
// depend on the global tile index and the candidate lists only, never on the shard, so the summation order -- and the
// result, bit for bit -- is the same for every partition of the grid.
// series length by floor(-8 log2(rho^2)) (see `expand`): filled by far_terms_table(), uploaded once per device
__constant__ unsigned char FAR_TERMS[256];

void far_terms_table(unsigned char *tab) {
    for (int t = 0; t < 256; t++) {
        const double lg = 0.5 * (t / 8.0 - 0.087);  // lower bound of log2(1 / rho) for this index
        int n = K1;
        if (lg > (double)SD_FAR_LOG2_RHO_INV) {
            n = (int)((double)K1 * (double)SD_FAR_LOG2_RHO_INV / lg + 1.02);
            if (n > K1) n = K1;
        }
        tab[t] = (unsigned char)n;
    }
}

// terms needed for ratio^2 = rho2: (n + 1) rho^n <= (K1 + 1) SD_FAR_RHO^K1 (the bound of the full series at the far
// criterion)  <=>  n >= ~K1 log2(1 / SD_FAR_RHO) / log2(1 / rho).  -log2(rho^2) is read off the exponent and the top
// mantissa bits of rho^2 in steps of 1/8 (a lower bound: the series is never shorter than the rule asks) and indexes
// a 256-entry table -- six integer instructions instead of ~30 with a float logarithm and division
__device__ __forceinline__ int far_terms(double rho2) {
    const int t8 = (0x3ff00000 - __double2hiint(rho2)) >> 17;  // floor(8 * -L), L <= log2(rho^2) <= L + 0.086
    return FAR_TERMS[min(max(t8, 0), 255)];
}

__global__ void __launch_bounds__(THREADS, 3) k_far_coeffs(LineArgs a, int lev, int count_stats, int nsplit, double *part) {
    __shared__ int s_ja[FAR_MAX_SRC], s_jb[FAR_MAX_SRC];
    extern __shared__ __align__(16) unsigned char far_smem[];
    FarRec *const s_rec = reinterpret_cast<FarRec *>(far_smem);                        // [FAR_CH] dense records of the chunk
    unsigned char *const s_mask = reinterpret_cast<unsigned char *>(s_rec + FAR_CH);   // [FAR_CH] their child masks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.y;
    const int shift = a.fg.tile_shift[lev], tile_px = a.fg.tile[lev];
    const bool has_parent = lev + 1 < a.n_act;
    const int plev = has_parent ? lev + 1 : lev;
    const int pshift = a.fg.tile_shift[plev];
    const int group = (a.far_tile0[lev] >> SD_FAR_SHIFT) + (int)blockIdx.x / nsplit;  // = parent tile index
    const int split = (int)blockIdx.x % nsplit;
    const int child0 = group << SD_FAR_SHIFT;
    const int tile_w = child0 + warp;
    const bool tile_ok = tile_w >= a.far_tile0[lev] && tile_w < a.far_tile0[lev] + a.far_ntl[lev];
    const int tile = tile_ok ? tile_w : a.far_tile0[lev];
    const int64_t t0 = (int64_t)tile * tile_px;
    const int64_t t1 = (t0 + tile_px < a.N) ? t0 + tile_px : a.N;
    const double nu_c = a.fg.geom[lev][3 * tile], h = a.fg.geom[lev][3 * tile + 1];
    const int ptile = group;
    const int64_t pt0 = (int64_t)ptile * a.fg.tile[plev];
    const int64_t pt1 = (pt0 + a.fg.tile[plev] < a.N) ? pt0 + a.fg.tile[plev] : a.N;
    // children of this group that belong to the launched range (bit c = tile child0 + c)
    unsigned valid = 0;
#pragma unroll
    for (int cc = 0; cc < WARPS; cc++) {
        const int tc = child0 + cc;
        if (tc >= a.far_tile0[lev] && tc < a.far_tile0[lev] + a.far_ntl[lev]) valid |= 1u << cc;
    }
    const size_t drow = (size_t)d * a.L;
    const int *list_d = a.cls_list + drow;
    const unsigned long long lmask = (1ull << a.fg.l_bits) - 1ull;  // line index = low bits of a window-edge key
    // candidate lists: source 3 m + kind (kind 0: centre range of class m; 1 / 2: window starts / ends of class m)
    const int n_src = 3 * (lev + 1);
    for (int src = warp; src < n_src; src += WARPS) {
        const int m = src / 3, kind = src - 3 * m;
        int ja = 0, jb = 0;
        if (!has_parent) {
            if (kind == 0) {
                ja = a.cls_off[d * (SD_NCLS + 1) + SD_FC0 + m];
                jb = a.cls_off[d * (SD_NCLS + 1) + SD_FC0 + m + 1];
            }
        } else if (kind == 0) {
            fc_centre_range(a, d, m, ((long long)ptile - 1) * a.fg.tile[plev], ((long long)ptile + 2) * a.fg.tile[plev], ja, jb);
        } else {
            fc_edge_range(a, d, kind - 1, m, pt0, pt1, ja, jb);
        }
        // this CTA's slice of the range (a function of the range and nsplit only)
        const int len = (jb - ja + nsplit - 1) / nsplit;
        const int sa = min(ja + split * len, jb), sb = min(sa + len, jb);
        if (lane == 0) { s_ja[src] = sa; s_jb[src] = sb; }
    }
    __syncthreads();
    double C[K1];
#pragma unroll
    for (int k = 0; k < K1; k++) C[k] = 0.0;
    unsigned long long n_far = 0, n_terms = 0;
    const unsigned lt_mask = (1u << lane) - 1u;

    // One accepted pair: Taylor coefficients of its two poles about the tile centre.  Called with a dense batch of
    // pairs (one per lane); `have` is false only in the last, partial batch.
    auto expand = [&](bool have, const FarRec &r) {
        double An = 0.0, w1r = 0.0, w1i = 0.0, w2r = 0.0, w2i = 0.0, v1i = 0.0, v2i = 0.0;
        int nterms = 0;
        if (have) {
            const double g = r.y * r.dw;                                // Lorentz half-width in Hz
            An = r.K * r.dw * (0.5 * sdm::INV_SQRT_PI);
            const double adw = 0.7071067811865476 * r.dw;
            // v = 1 / (D - i g) = (D + i g) / (D^2 + g^2),  w = -h v  for the two poles
            const double D1 = nu_c - (r.nu + adw), D2 = nu_c - (r.nu - adw);
            const double q1 = sdm::rcp_fast(fma(D1, D1, g * g)), q2 = sdm::rcp_fast(fma(D2, D2, g * g));
            v1i = g * q1; v2i = g * q2;
            const double i1 = -h * q1, i2 = -h * q2;
            w1r = D1 * i1; w1i = g * i1; w2r = D2 * i2; w2i = g * i2;
            nterms = far_terms(h * h * fmax(q1, q2));
            if (count_stats) n_far++;
        }
        // queue neighbours are neighbours in frequency, at similar distances from the tile: warp-uniform series length
        const int nt = __reduce_max_sync(0xffffffffu, nterms);
        if (count_stats && have) n_terms += (unsigned long long)min(K1, 4 * ((nt + 3) / 4));  // terms the loop below executes
        // Im(v w^k) by the real three-term recurrence of a geometric sequence of complex numbers,
        //   s_(k+1) = 2 Re(w) s_k - |w|^2 s_(k-1),  s_0 = Im v, s_(-1) = Im(v / w) = Im(-1 / h) = 0,
        // two instructions per pole and term instead of the four of a complex product (the recurrence loses about one
        // bit per step relative to |w|^k, i.e. < 1e-13 over the series).  No division by h: a one-pixel tile (h = 0)
        // gets its exact constant term.
        const double a1 = w1r + w1r, b1 = fma(w1r, w1r, w1i * w1i), a2 = w2r + w2r, b2 = fma(w2r, w2r, w2i * w2i);
        double s1 = v1i, s1p = 0.0, s2 = v2i, s2p = 0.0;
#pragma unroll
        for (int k = 0; k < K1; k++) {
            if (k % 4 == 0 && k >= nt) break;  // checked every fourth term (the extra terms only add accuracy)
            C[k] = fma(An, s1 + s2, C[k]);
            if (k + 1 < K1) {
                double t;
                t = fma(a1, s1, -(b1 * s1p)); s1p = s1; s1 = t;
                t = fma(a2, s2, -(b2 * s2p)); s2p = s2; s2 = t;
            }
        }
    };

    // Per chunk of FAR_CH candidates:
    //   test     a thread gathers the window record of its candidates ONCE and tests it against all eight children;
    //   compact  the candidates wanted by ANY child (of the whole group, launched or not: the list must not depend on
    //            the shard) are packed densely, in list order, into shared memory: record (cp.async) + child mask;
    //   expand   every warp walks the dense list 32 entries at a time, one pair per lane, skipping the entries whose
    //            mask lacks its child.
    constexpr int ROUNDS = FAR_CH / THREADS;
    constexpr int SEGS = ROUNDS * WARPS;       // (round, warp) segments of 32 candidates, in list order
    static_assert(SEGS == 32, "one lane per segment in the offset scan");
    __shared__ int s_cnt[SEGS];
    // one candidate stream: the slices of all sources, concatenated in source order
    __shared__ int s_pre[FAR_MAX_SRC + 1];
    if (tid == 0) {
        int acc = 0;
        for (int src = 0; src < FAR_MAX_SRC; src++) {
            s_pre[src] = acc;
            if (src < n_src) acc += s_jb[src] - s_ja[src];
        }
        s_pre[FAR_MAX_SRC] = acc;
    }
    __syncthreads();
    const int total = s_pre[FAR_MAX_SRC];
    if (total == 0) {  // nothing to expand for this group (CTA-uniform): zeros
        if (tile_ok) {
            const size_t tl = (size_t)d * a.far_ntl[lev] + (tile - a.far_tile0[lev]);
            if (nsplit > 1) part[(tl * nsplit + split) * K1 + lane] = 0.0;
            else a.far_coef[lev][tl * K1 + lane] = 0.0;
        }
        return;
    }
    {
        for (int base = 0; base < total; base += FAR_CH) {
            // ---- test
            unsigned mk[ROUNDS];
            int ll[ROUNDS], rk[ROUNDS];
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const int v = base + r * THREADS + tid;
                unsigned mask = 0;
                int l = 0;
                if (v < total) {
                    int src = 0, off = 0;
#pragma unroll
                    for (int q = 1; q < FAR_MAX_SRC; q++) {
                        const int pq = s_pre[q];   // non-decreasing
                        if (v >= pq) { src = q; off = pq; }
                    }
                    const int j = s_ja[src] + (v - off);
                    const int kind = src % 3;
                    l = (kind == 0) ? list_d[j] : (int)(a.fg.edge_keys[j] & lmask);
                    const PairWin pw = load_win(a.win + drow + l);
                    bool okp;
                    if (kind == 0) {  // (A) / top level: served here unless the multipole path takes the pair
                        okp = !((pw.sat >> lev) & 1u);
                    } else {          // (B): far from the parent by distance; both edges inside the parent: via its start
                        const int dsp = (pw.cpix >> pshift) - ptile;
                        okp = (dsp >= 2 || dsp <= -2) && ((kind == 1 ? pw.lo : pw.hi) > pt0) && !(kind == 2 && pw.lo > pt0 && pw.lo < pt1);
                    }
                    if (okp) {
#pragma unroll
                        for (int cc = 0; cc < WARPS; cc++) {
                            const int tcc = child0 + cc;
                            const int64_t c0 = (int64_t)tcc * tile_px;
                            const int64_t c1 = (c0 + tile_px < a.N) ? c0 + tile_px : a.N;
                            if (tcc < a.fg.n_tiles[lev] && pair_is_far(pw, lev, shift, tcc, c0, c1)) mask |= 1u << cc;
                        }
                    }
                }
                const unsigned nz = __ballot_sync(0xffffffffu, mask != 0);
                mk[r] = mask; ll[r] = l; rk[r] = __popc(nz & lt_mask);
                if (lane == 0) s_cnt[r * WARPS + warp] = __popc(nz);
            }
            __syncthreads();
            // ---- compact: exclusive offsets of the (round, warp) segments, one segment per lane
            const int cnt_l = s_cnt[lane];
            int incl = cnt_l;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o2);
                if (lane >= o2) incl += t;
            }
            const int n_dense = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const int seg = r * WARPS + warp;
                const int off = __shfl_sync(0xffffffffu, incl - cnt_l, seg);
                if (mk[r]) {
                    const int pos = off + rk[r];
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(s_rec + pos);
                    const LineRec *srcp = a.rec + drow + ll[r];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(srcp) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16u),
                                 "l"(reinterpret_cast<const char *>(srcp) + 16) : "memory");
                    s_mask[pos] = (unsigned char)(mk[r] & valid);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            // ---- expand: this warp's child
            for (int i0 = 0; i0 < n_dense && tile_ok; i0 += 32) {
                const int i = i0 + lane;
                const bool have = (i < n_dense) && ((s_mask[i] >> warp) & 1u);
                if (!__any_sync(0xffffffffu, have)) continue;
                expand(have, s_rec[i < n_dense ? i : 0]);
            }
            __syncthreads();  // the chunk buffers are overwritten by the next chunk
        }
    }
    // deterministic reduction over the lanes through shared memory (the chunk buffers are free now): two halves of 16
    // coefficients, lane k sums coefficient k of the 32 lanes in lane order -- a third of the instructions of 32
    // butterfly reductions
    static_assert(FAR_SMEM >= sizeof(double) * WARPS * (K1 / 2) * 33, "reduction scratch fits the chunk buffers");
    double (*red)[33] = reinterpret_cast<double (*)[33]>(far_smem) + warp * (K1 / 2);
    const size_t tl = (size_t)d * a.far_ntl[lev] + (tile - a.far_tile0[lev]);
#pragma unroll
    for (int half = 0; half < 2; half++) {
#pragma unroll
        for (int kk = 0; kk < K1 / 2; kk++) red[kk][lane] = C[half * (K1 / 2) + kk];
        __syncwarp();
        if (lane < K1 / 2) {
            double v = 0.0;
#pragma unroll 8
            for (int i = 0; i < 32; i++) v += red[lane][i];
            if (tile_ok) {
                const int k = half * (K1 / 2) + lane;
                if (nsplit > 1) part[(tl * nsplit + split) * K1 + k] = v;
                else a.far_coef[lev][tl * K1 + k] = v;
            }
        }
        __syncwarp();
    }
    if (count_stats) {  // every far pair stands for one region-I evaluation per tile pixel inside the shard
        for (int o2 = 16; o2; o2 >>= 1) {
            n_far += __shfl_xor_sync(0xffffffffu, n_far, o2);
            n_terms += __shfl_xor_sync(0xffffffffu, n_terms, o2);
        }
        const int64_t e0 = t0 > a.p0 ? t0 : a.p0, e1 = t1 < a.p1 ? t1 : a.p1;
        if (lane == 0 && n_far) {
            if (e1 > e0) atomicAdd(&a.stats[8], n_far * (unsigned long long)(e1 - e0));  // evaluations these expansions replace
            atomicAdd(&a.stats[9], n_far);     // executed expansions (pair, tile)
            atomicAdd(&a.stats[10], n_terms);  // executed series terms (both poles count as one)
        }
    }
}

// slices per level (constants: the summation order must not depend on the shard)
__host__ __device__ constexpr int far_nsplit_base(int lev, bool top) { return top ? 32 : (lev <= 1 ? 1 : 8); }
// SD_FAR_NSPLIT_SCALE (tuning experiments only: it changes the grouping of the partial sums, i.e. the last bits)
static int far_nsplit(int lev, bool top) {
    static const int scale = env_int_early("SD_FAR_NSPLIT_SCALE", 1);
    return far_nsplit_base(lev, top) * (scale >= 1 && scale <= 8 ? scale : 1);
}

// sum of the nsplit partial coefficient sets of a level, in slice order
__global__ void k_far_reduce(int n, int nsplit, const double *__restrict__ part, double *__restrict__ coef) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (depth, tile) * K1 + k
    if (i >= n) return;
    const int td = i / K1, k = i - td * K1;
    double v = 0.0;
    for (int s2 = 0; s2 < nsplit; s2++) v += part[((size_t)td * nsplit + s2) * K1 + k];
    coef[i] = v;
}

// ------------------------------------------------------------------------------------------------------------------
// Multipole moments:  M_k(s) = sum_pairs A Im(u+^k + u-^k),  k = 1..K1,  u = (pole - c_s) / scale_s,  over the pairs
// centred in tile s that are SATURATED at the tile's level (window covering the whole interaction neighbourhood), so that
//   sum_pairs contribution(nu) = (1 / scale_s) sum_k M_k / tau^(k+1),  tau = (nu - c_s) / scale_s,
// for every pixel at least two tiles away.  Every pair is expanded ONCE, about its tile of level lmin (k_s2m), into the
// bucket of its highest saturated level hs (saturation is monotone: saturated at a level => at every lower level >=
// lmin); k_m2m translates the buckets to the parent tiles level by level (exact: a finite binomial sum) and adds up, per
// level k, the buckets hs >= k -- the moments k_m2l needs there.  (The direct scheme of round 1 expanded a whole-grid
// pair about ~21 target tiles per level.)
//
// k_s2m(m): one warp per run of eight level-m tiles (the children of one level-(m+1) tile) and depth.  The pairs of the
// run are one contiguous piece of the class list of lmin = m (centres descend along it).  Per batch of 32 pairs every
// lane expands ITS pair about the pair's own tile (three-term recurrence for Im(u^k)) and writes the 32 moments to
// shared memory; then lane k walks the batch in list order and adds moment k of every pair to the accumulator of the
// pair's bucket, flushing to global memory whenever the tile changes (tiles without pairs get zeros).  All lanes busy,
// fixed order, nothing depends on the shard.
constexpr int S2M_WARPS = 4;

// pixels of the shard inside the interaction list of level-`lev` tile s (statistics: evaluations one saturated pair stands for)
__device__ __forceinline__ long long il_pixels(const LineArgs &a, int lev, int s) {
    const long long T = a.fg.tile[lev];
    long long nb0 = 0, nb1 = a.N;
    if (lev + 1 < a.n_act) {
        const long long P = s >> SD_FAR_SHIFT, Tp = a.fg.tile[lev + 1];
        nb0 = (P - 1) * Tp > 0 ? (P - 1) * Tp : 0;
        nb1 = (P + 2) * Tp < a.N ? (P + 2) * Tp : a.N;
    }
    const long long nr0 = ((long long)s - 1) * T > 0 ? ((long long)s - 1) * T : 0;
    const long long nr1 = ((long long)s + 2) * T < a.N ? ((long long)s + 2) * T : a.N;
    auto clip = [&](long long x0, long long x1) {
        const long long e0 = x0 > a.p0 ? x0 : a.p0, e1 = x1 < a.p1 ? x1 : a.p1;
        return e1 > e0 ? e1 - e0 : 0LL;
    };
    return clip(nb0, nb1) - clip(nr0, nr1);
}

__global__ void __launch_bounds__(32 * S2M_WARPS) k_s2m(LineArgs a, int m, int count_stats) {
    __shared__ double s_buf[S2M_WARPS][K1][33];
    __shared__ int s_tile[S2M_WARPS][32];
    __shared__ int s_hs[S2M_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.y;
    const int nt = a.fg.n_tiles[m];
    const int run = (int)blockIdx.x * S2M_WARPS + warp;
    const int t_lo = run << SD_FAR_SHIFT;
    if (t_lo >= nt) return;  // warps are independent: no CTA barrier below
    const int t_hi = min(t_lo + (1 << SD_FAR_SHIFT), nt) - 1;
    const int shift = a.fg.tile_shift[m];
    const long long T = a.fg.tile[m];
    const double *__restrict__ gm = a.fg.geom[m];
    const size_t drow = (size_t)d * a.L;
    const int *list_d = a.cls_list + drow;
    double (*buf)[33] = s_buf[warp];
    int *const b_tile = s_tile[warp], *const b_hs = s_hs[warp];
    const int nb = a.n_act - m;  // buckets hs = m .. n_act - 1
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int cur = -1;              // tile whose sums are in the accumulators (-1: none)
    int next_out = t_hi;       // highest tile of the run not written yet
    unsigned long long n_pix = 0, n_exp = 0;
    auto store_tile = [&](int tile, double v0, double v1, double v2, double v3) {
        const size_t o = ((size_t)d * nt + tile) * K1 + lane;
        a.far_bkt[m][m][o] = v0;
        if (nb > 1) a.far_bkt[m][m + 1][o] = v1;
        if (nb > 2) a.far_bkt[m][m + 2][o] = v2;
        if (nb > 3) a.far_bkt[m][m + 3][o] = v3;
    };
    auto close_down_to = [&](int tile) {  // write the open tile and zeros for the untouched ones above `tile`
        if (cur >= 0) {
            store_tile(cur, acc0, acc1, acc2, acc3);
            next_out = cur - 1;
            acc0 = acc1 = acc2 = acc3 = 0.0;
            cur = -1;
        }
        for (; next_out > tile; next_out--) store_tile(next_out, 0.0, 0.0, 0.0, 0.0);
    };
    int ja, jb;
    fc_centre_range(a, d, m, (long long)t_lo * T, ((long long)t_hi + 1) * T, ja, jb);
    for (int j0 = ja; j0 < jb; j0 += 32) {
        const int j = j0 + lane;
        int hs = -1, tile = 0;
        double An = 0.0, u1r = 0.0, u2r = 0.0, ui = 0.0;
        if (j < jb) {
            const int l = list_d[j];
            const PairWin pw = load_win(a.win + drow + l);
            tile = pw.cpix >> shift;
            hs = pw.sat ? 31 - __clz((unsigned)pw.sat) : -1;
            if (hs >= m) {
                const double2 *rp = reinterpret_cast<const double2 *>(a.rec + drow + l);
                const double2 r0 = __ldg(rp), r1 = __ldg(rp + 1);   // nu, dw, y, K
                const double c_s = gm[3 * tile], sc = gm[3 * tile + 2];
                const double inv_sc = sc > 0.0 ? 1.0 / sc : 0.0;
                const double g = r1.x * r0.y, adw = 0.7071067811865476 * r0.y;
                An = r1.y * r0.y * (0.5 * sdm::INV_SQRT_PI);
                const double dc = r0.x - c_s;
                u1r = (dc + adw) * inv_sc; u2r = (dc - adw) * inv_sc; ui = g * inv_sc;
                if (count_stats) {
                    for (int lv = m; lv <= hs; lv++) n_pix += (unsigned long long)il_pixels(a, lv, pw.cpix >> a.fg.tile_shift[lv]);
                    n_exp += (unsigned long long)(hs - m + 1);
                }
            }
        }
        b_tile[lane] = tile;
        b_hs[lane] = hs;
        const double a1 = u1r + u1r, b1 = fma(u1r, u1r, ui * ui), a2 = u2r + u2r, b2 = fma(u2r, u2r, ui * ui);
        double s1 = ui, s1p = 0.0, s2 = ui, s2p = 0.0;   // Im(u^1), Im(u^0)
#pragma unroll
        for (int k = 0; k < K1; k++) {
            buf[k][lane] = An * (s1 + s2);
            if (k + 1 < K1) {
                double t;
                t = fma(a1, s1, -(b1 * s1p)); s1p = s1; s1 = t;
                t = fma(a2, s2, -(b2 * s2p)); s2p = s2; s2 = t;
            }
        }
        __syncwarp();
        const int nin = min(32, jb - j0);
        for (int i = 0; i < nin; i++) {
            const int h = b_hs[i] - m;   // warp-uniform
            if (h < 0) continue;
            const int ti = b_tile[i];
            if (ti != cur) {
                close_down_to(ti);
                cur = ti;
            }
            const double v = buf[lane][i];
            if (h == 0) acc0 += v;
            else if (h == 1) acc1 += v;
            else if (h == 2) acc2 += v;
            else acc3 += v;
        }
        __syncwarp();
    }
    close_down_to(t_lo - 1);
    if (count_stats) {
        for (int o2 = 16; o2; o2 >>= 1) {
            n_pix += __shfl_xor_sync(0xffffffffu, n_pix, o2);
            n_exp += __shfl_xor_sync(0xffffffffu, n_exp, o2);
        }
        if (lane == 0 && n_exp) {
            atomicAdd(&a.stats[8], n_pix);   // evaluations the multipole path replaces
            atomicAdd(&a.stats[12], n_exp);  // (pair, level) products served by multipole moments
        }
    }
}

// k_m2m(lev): one warp per (level-`lev` parent tile P, depth), lane k = moment k.  For each child s of P at level lev - 1:
//   * far_mom[lev-1][s] = sum_{h >= lev-1} far_bkt[lev-1][h][s]  (what k_m2l uses at the children's level);
//   * far_bkt[lev][h][P] += translate(far_bkt[lev-1][h][s]) for h >= lev:
//       M'_k = sum_{j=1..k} C(k, j) r^j delta^(k-j) M_j,   r = scale_s / scale_P,  delta = (c_s - c_P) / scale_P
//     (children in index order; the parent's own pairs with lmin = lev were written by k_s2m(lev) before).
// lev == n_act: only the first step, for the top level.
__constant__ double M2M_INV[K1 + 2];  // 1 / j

__global__ void __launch_bounds__(128) k_m2m(LineArgs a, int lev) {
    __shared__ double s_m[4][3][K1 + 1];   // child moments M_1..M_K1 of up to three buckets (index j; [0] unused = 0)
    __shared__ double s_dp[4][K1 + 1];     // delta^i, i = 0..K1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.y;
    const int cl = lev - 1;                          // children's level
    const int ntc = a.fg.n_tiles[cl];
    const int P = (int)blockIdx.x * 4 + warp;
    if ((P << SD_FAR_SHIFT) >= ntc) return;
    const bool translate = lev < a.n_act;
    const int ntp = translate ? a.fg.n_tiles[lev] : 0;
    const int nh = a.n_act - lev;                    // buckets h = lev .. n_act - 1 go up (<= 3)
    const int kk = lane + 1;                         // this lane's moment index (power)
    double up0 = 0.0, up1 = 0.0, up2 = 0.0;
    double c_P = 0.0, inv_sP = 0.0;
    if (translate) {
        c_P = a.fg.geom[lev][3 * P];
        const double sP = a.fg.geom[lev][3 * P + 2];
        inv_sP = sP > 0.0 ? 1.0 / sP : 0.0;
    }
    double (*mm)[K1 + 1] = s_m[warp];
    double *dp = s_dp[warp];
    for (int c = 0; c < (1 << SD_FAR_SHIFT); c++) {
        const int s = (P << SD_FAR_SHIFT) + c;
        if (s >= ntc) break;
        const size_t o = ((size_t)d * ntc + s) * K1 + lane;
        double tot = a.far_bkt[cl][cl][o];
        double b0 = 0.0, b1 = 0.0, b2 = 0.0;
        if (nh > 0) { b0 = a.far_bkt[cl][lev][o]; tot += b0; }
        if (nh > 1) { b1 = a.far_bkt[cl][lev + 1][o]; tot += b1; }
        if (nh > 2) { b2 = a.far_bkt[cl][lev + 2][o]; tot += b2; }
        a.far_mom[cl][o] = tot;
        if (!translate) continue;
        const unsigned any = __ballot_sync(0xffffffffu, b0 != 0.0 || b1 != 0.0 || b2 != 0.0);
        if (any == 0) continue;
        const double r = a.fg.geom[cl][3 * s + 2] * inv_sP, delta = (a.fg.geom[cl][3 * s] - c_P) * inv_sP;
        __syncwarp();
        mm[0][kk] = b0; mm[1][kk] = b1; mm[2][kk] = b2;
        {   // delta^i by binary powering: lane i -> delta^i (i = 0..31), lane 0 also writes delta^32
            double pw = 1.0, base = delta;
#pragma unroll
            for (int bit = 0; bit < 5; bit++) {
                if ((lane >> bit) & 1) pw *= base;
                base *= base;
            }
            dp[lane] = pw;
            if (lane == 0) dp[K1] = base;   // delta^32 after five squarings
        }
        __syncwarp();
        // row kk: coefficient of M_j is C(kk, j) r^j delta^(kk - j); walk j upwards with the running C(kk, j) r^j
        double cf = (double)kk * r;   // j = 1
        double rem = (double)(kk - 1);   // kk - j as a double (integer-to-double conversions are slow)
        for (int j = 1; j <= K1; j++) {
            if (j <= kk) {
                const double w = cf * dp[kk - j];
                up0 = fma(w, mm[0][j], up0);
                up1 = fma(w, mm[1][j], up1);
                up2 = fma(w, mm[2][j], up2);
                cf *= (r * M2M_INV[j + 1]) * rem;
                rem -= 1.0;
            }
        }
    }
    if (translate && P < ntp) {
        const size_t o = ((size_t)d * ntp + P) * K1 + lane;
        if (nh > 0) a.far_bkt[lev][lev][o] += up0;
        if (nh > 1) a.far_bkt[lev][lev + 1][o] += up1;
        if (nh > 2) a.far_bkt[lev][lev + 2][o] += up2;
    }
}

// Tile-to-tile translation of the multipole moments into Taylor coefficients, added to the direct part:
//   L_n(t) += sum_s (b^n / d) sum_k C(n + k, n) a^k M_k(s),   d = c_t - c_s, a = scale_s / d, b = -h_t / d,
// over the source tiles s of the interaction list of t: |s - t| >= 2 and parent(s) within one tile of parent(t) (all
// tiles with |s - t| >= 2 at the top level).  The matrix is real and independent of the depth point: a CTA takes the
// eight children of a parent (warp w = target tile) and M2L_DC depth points, lane n owns coefficient n, builds its
// matrix row on the fly (ratio table in shared memory) and multiplies it with the moments of the chunk's depth points,
// staged in shared memory (every lane reads the same moment: broadcast).  The k-sum stops where the multipole series
// has converged for this tile distance (same rule as the direct expansion).  Fixed source order: shard-invariant.
template <int M2L_DC>
__global__ void __launch_bounds__(THREADS) k_m2l(LineArgs a, int lev, int count_stats) {
    __shared__ double s_R[K1][K1];                           // s_R[k][n] = (n + k + 2) / (k + 2)
    __shared__ __align__(16) double s_M[2][K1][M2L_DC + 2];   // moments of the staged source tile, [k][depth]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool top = lev + 1 >= a.n_act;
    const int nt = a.fg.n_tiles[lev];
    const int P = (a.far_tile0[lev] >> SD_FAR_SHIFT) + (int)blockIdx.x;
    const int t = (P << SD_FAR_SHIFT) + warp;
    const bool tile_ok = t >= a.far_tile0[lev] && t < a.far_tile0[lev] + a.far_ntl[lev];
    const int d0 = (int)blockIdx.y * M2L_DC;
    const int nd = min(M2L_DC, a.D - d0);
    for (int i = tid; i < K1 * K1; i += THREADS) {
        const int k = i / K1, n = i - k * K1;
        s_R[k][n] = (double)(n + k + 2) / (double)(k + 2);
    }
    const int s_lo = top ? 0 : max(((P - 1) << SD_FAR_SHIFT), 0);
    const int s_hi = top ? nt : min(((P + 2) << SD_FAR_SHIFT), nt);
    const double *__restrict__ gm = a.fg.geom[lev];
    const double c_t = gm[3 * (tile_ok ? t : 0)], h_t = gm[3 * (tile_ok ? t : 0) + 1];
    const double *__restrict__ mom = a.far_mom[lev];
    const int src0 = 0, nsrc = nt;
    // staging of a source tile's moments in two steps, so that the global loads of the NEXT tile are in flight while
    // the current one is multiplied: fetch() issues the loads into registers, commit() writes them to shared memory
    constexpr int NST = (M2L_DC * K1 + THREADS - 1) / THREADS;
    double pre[NST];
    auto fetch = [&](int s) {
#pragma unroll
        for (int r = 0; r < NST; r++) {
            const int i = tid + r * THREADS;
            const int dd = i / K1, k = i - dd * K1;
            pre[r] = (i < M2L_DC * K1 && dd < nd) ? __ldg(mom + ((size_t)(d0 + dd) * nsrc + (s - src0)) * K1 + k) : 0.0;
        }
    };
    auto commit = [&](int buf) {  // returns whether this thread staged a non-zero moment
        int nz = 0;
#pragma unroll
        for (int r = 0; r < NST; r++) {
            const int i = tid + r * THREADS;
            const int dd = i / K1, k = i - dd * K1;
            if (i < M2L_DC * K1) s_M[buf][k][dd] = pre[r];
            nz |= (pre[r] != 0.0);
        }
        return nz;
    };
    double acc[M2L_DC];
#pragma unroll
    for (int dd = 0; dd < M2L_DC; dd++) acc[dd] = 0.0;
    unsigned long long n_m2l = 0;
    int nz0 = 0;
    if (s_lo < s_hi) { fetch(s_lo); nz0 = commit(0); }
    int nz_cur = __syncthreads_or(nz0);   // source tiles without saturated pairs (all moments zero) are skipped
    for (int s = s_lo; s < s_hi; s++) {
        const int buf = (s - s_lo) & 1;
        if (s + 1 < s_hi) fetch(s + 1);
        const int ds = s - t;
        if (nz_cur && tile_ok && (ds >= 2 || ds <= -2)) {
            const double c_s = gm[3 * s], sc = gm[3 * s + 2];
            const double dd_ = c_t - c_s, inv_d = 1.0 / dd_;
            const double aa = sc * inv_d, bb = -h_t * inv_d;
            // b^n by binary powering (lane n)
            double f = inv_d, base = bb;
#pragma unroll
            for (int bit = 0; bit < 5; bit++) {
                if ((lane >> bit) & 1) f *= base;
                base *= base;
            }
            // the multipole series in (pole - c_s) / (nu - c_s): poles up to 1 + overhang scale lengths from c_s,
            // pixels of t at least |d| - h_t away
            const double rr = (1.0 + SD_FAR_OVERHANG) * sc / (fabs(dd_) - h_t);
            const int kmax = min(K1, 4 * ((far_terms(rr * rr) + 3) / 4));
            double coef = f * (double)(lane + 1) * aa;
            for (int k = 0; k < kmax; k++) {
                const double *__restrict__ mk = s_M[buf][k];
#pragma unroll
                for (int dd = 0; dd < M2L_DC; dd += 2) {
                    const double2 mv = *reinterpret_cast<const double2 *>(mk + dd);
                    acc[dd] = fma(coef, mv.x, acc[dd]);
                    acc[dd + 1] = fma(coef, mv.y, acc[dd + 1]);
                }
                coef *= aa * s_R[k][lane];
            }
            if (count_stats) n_m2l += (unsigned long long)nd * kmax;
        }
        int nz_next = 0;
        if (s + 1 < s_hi) nz_next = commit(buf ^ 1);
        nz_cur = __syncthreads_or(nz_next);
    }
    if (tile_ok) {
#pragma unroll
        for (int dd = 0; dd < M2L_DC; dd++)
            if (dd < nd) a.far_coef[lev][((size_t)(d0 + dd) * a.far_ntl[lev] + (t - a.far_tile0[lev])) * K1 + lane] += acc[dd];
    }
    if (count_stats && lane == 0 && n_m2l) atomicAdd(&a.stats[13], n_m2l);  // executed (source, target, depth, k) row steps per lane
}

// ------------------------------------------------------------------------------------------------------------------
template <int P, int NW, bool STATS, int RCP, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) k_lines(LineArgs a) {
    constexpr int TILE = 32 * NW * P;
    constexpr int SPAN = 32 * P;
    constexpr int SUB = 1 << SD_FAR_TILE0_SHIFT;        // level-0 far-field tile: 64 pixels = 2 register slots
    constexpr int NSUB = (SPAN + SUB - 1) / SUB;        // level-0 tiles per warp span (4 for P = 8)
    constexpr int SLOTS_PER_SUB = SUB / 32;
    constexpr unsigned ALL_SUB = (1u << NSUB) - 1u;
    __shared__ WEntry s_ent[NW][32];  // every warp streams its own batches: no CTA barrier in the main loop
    __shared__ unsigned char s_msk[NW][32];  // per entry: level-0 tiles of the span that are evaluated directly
    __shared__ double s_acc[NW][SPAN];  // accumulators of the pixel-parallel mixed path
    __shared__ double s_nsub[NW][2 * NSUB];  // frequencies of the first / last pixel of every level-0 tile of the span
    // classes 0..6, then per lmin class m: far-capable pairs centred near the tile, their window starts / ends
    constexpr int NSRC = SD_FC0 + 3 * SD_FAR_LEVELS;
    __shared__ int s_ja[NSRC], s_jb[NSRC];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int d = blockIdx.y;
    const int64_t N = a.N, p0 = a.p0, p1 = a.p1, L = a.L;
    const int tile = a.tile0 + blockIdx.x;
    const int64_t t0 = (int64_t)tile * TILE;                       // GLOBAL tile [t0, t1)
    const int64_t t1 = (t0 + TILE < N) ? t0 + TILE : N;
    const int64_t ws = t0 + (int64_t)warp * SPAN;                  // this warp's global span [ws, we)
    const int64_t we = (ws + SPAN < t1) ? ws + SPAN : t1;
    const size_t drow = (size_t)d * L;
    const int *list_d = a.cls_list + drow;
    const double *__restrict__ nus = a.nus;
    constexpr bool FARCAP = (P == 8);  // the far field needs whole level-0 tiles per pair of register slots
    const bool use_far = FARCAP && a.fg.enabled != 0 && a.n_act > 0;
    const unsigned long long lmask = (1ull << a.fg.l_bits) - 1ull;  // line index = low bits of a window-edge key

    // pixel frequencies and accumulators live in registers for the whole kernel
    double nu_i[P], acc[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        nu_i[p] = nus[pix < N ? pix : N - 1];
        acc[p] = 0.0;
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
        s_acc[warp][p * 32 + lane] = 0.0;
    }
    if (lane < 2 * NSUB) {
        const int64_t e0 = ws + (int64_t)(lane >> 1) * SUB;
        int64_t pix = (lane & 1) ? ((e0 + SUB < we) ? e0 + SUB : we) - 1 : e0;
        pix = pix < ws ? ws : pix;
        s_nsub[warp][lane] = nus[pix < N ? pix : N - 1];
    }

    // candidate ranges of all sources, distributed over the warps of the CTA
    for (int src = warp; src < NSRC; src += NW) {
        int ja = 0, jb = 0;
        if (src < SD_FC0) class_range(a, d, src, t0, t1, ja, jb);
        else if (use_far) {
            const int m = (src - SD_FC0) / 3, kind = (src - SD_FC0) - 3 * m;
            if (m < a.n_act) {
                const int sh = a.fg.tile_shift[m];
                const long long ta = t0 >> sh, tb = (t1 - 1) >> sh;   // level-m tiles that overlap the CTA tile
                if (kind == 0) fc_centre_range(a, d, m, (ta - 1) << sh, (tb + 2) << sh, ja, jb);
                else {
                    const long long e1 = ((tb + 1) << sh) < N ? ((tb + 1) << sh) : N;
                    fc_edge_range(a, d, kind - 1, m, ta << sh, e1, ja, jb);
                }
            }
        }
        if (lane == 0) { s_ja[src] = ja; s_jb[src] = jb; }
    }
    __syncthreads();

    unsigned long long h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    unsigned nvalid_sub = 0;  // this lane's pixels that belong to the shard, per level-0 tile of the span, 8 bits each (statistics only)
#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        if ((pix < t1) && (pix >= p0) && (pix < p1)) nvalid_sub += 1u << (8 * (p / SLOTS_PER_SUB));
    }

    WEntry *const my = s_ent[warp];
    unsigned char *const my_msk = s_msk[warp];
    const bool warp_has_pixels = (ws < t1) && (ws < p1) && (we > p0);

    // far-wing (region I) evaluation of one staged entry for the pixels of this lane in the level-0 tiles of `msk`
    auto far_eval = [&](const double xl, const double inv_dw, const double eb, const double ec, const double Kc,
                        const double Kf, const unsigned msk) {
        if (msk == ALL_SUB) {
#pragma unroll
            for (int p = 0; p < P; p++) {
                double x = fma(nu_i[p], inv_dw, -xl);
                double q = x * x;
                double den = fma(q, q + eb, ec);
                double num = fma(Kf, q, Kc);
                acc[p] = fma(num, RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den), acc[p]);
            }
        } else {
#pragma unroll
            for (int p = 0; p < P; p++) {
                if (!((msk >> (p / SLOTS_PER_SUB)) & 1u)) continue;  // warp-uniform
                double x = fma(nu_i[p], inv_dw, -xl);
                double q = x * x;
                double den = fma(q, q + eb, ec);
                double num = fma(Kf, q, Kc);
                acc[p] = fma(num, RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den), acc[p]);
            }
        }
    };

    for (int src = 0; src < NSRC && warp_has_pixels; src++) {
        const int ja = s_ja[src], jb = s_jb[src];
        const int skind = src < SD_FC0 ? -1 : (src - SD_FC0) % 3;  // -1: half-width class, 0: centre list, 1 / 2: edge lists
        for (int base = ja; base < jb; base += 32) {
            {
            // ---- test the 32 candidates of this batch against THIS WARP's span ------------------------
            const int j = base + lane;
            bool pass = false;
            size_t o = 0;
            int lo = 0, hi = 0;
            unsigned msk = ALL_SUB;
            if (j < jb) {
                int l;
                if (src == 0) l = j;                                   // class 0: the nu-sorted line list itself
                else if (skind <= 0) l = list_d[j];                    // class lists
                else l = (int)(a.fg.edge_keys[j] & lmask);             // far-capable pairs with an edge inside the tile
                o = drow + l;
                const PairWin pw = load_win(a.win + o);
                lo = pw.lo;
                hi = pw.hi;
                pass = (lo < we) && (hi > ws) && (hi > lo);
                if (src == 0) pass = pass && (pw.cls == 0);
                else if (skind >= 0) {
                    // far-capable pair of class m = lmin: a level-0 tile of the span is evaluated here unless the pair
                    // is far from the level-m tile that holds it (then the polynomials carry the contribution)
                    const int m = pw.lmin, sh = a.fg.tile_shift[m], sj = pw.cpix >> sh;
                    if (skind > 0) {
                        // edge lists: only pairs the centre list of this CTA does not already deliver; both edges
                        // inside the level-m tile range: via the start
                        const int ta = (int)(t0 >> sh), tb = (int)((t1 - 1) >> sh);
                        const int64_t r0 = (int64_t)ta << sh;
                        const int64_t r1 = (((int64_t)tb + 1) << sh) < N ? (((int64_t)tb + 1) << sh) : N;
                        pass = pass && (sj < ta - 1 || sj > tb + 1) && ((skind == 1 ? lo : hi) > r0) && !(skind == 2 && lo > r0 && lo < r1);
                    }
                    msk = 0;
                    if (m == 0) {
#pragma unroll
                        for (int i = 0; i < NSUB; i++) {
                            const int64_t c0 = ws + (int64_t)i * SUB;
                            const int64_t c1 = (c0 + SUB < N) ? c0 + SUB : N;
                            if (c0 < we && !pair_is_far(pw, 0, SD_FAR_TILE0_SHIFT, (int)(c0 >> SD_FAR_TILE0_SHIFT), c0, c1)) msk |= 1u << i;
                        }
                    } else {
                        const int tm = (int)(ws >> sh);
                        const int64_t c0 = (int64_t)tm << sh;
                        const int64_t c1 = (c0 + ((int64_t)1 << sh) < N) ? c0 + ((int64_t)1 << sh) : N;
                        if (!pair_is_far(pw, m, sh, tm, c0, c1)) msk = ALL_SUB;
                    }
                    pass = pass && (msk != 0);
                }
            }
            if (!__any_sync(0xffffffffu, pass)) continue;
            // ---- stage: hoist the per-(line, depth) constants; far-wing entries first, mixed ones from the back
            WEntry e;
            bool ff = false;
            if (pass) {
                const LineRec r = a.rec[o];
                const double yy = r.y * r.y;
                const double c1 = yy + 0.5;
                e.xl = r.nu * r.inv_dw;
                e.inv_dw = r.inv_dw;
                e.b = 2.0 * yy - 1.0;
                e.c = c1 * c1;
                e.Kf = r.K * r.y * sdm::INV_SQRT_PI;
                e.Kc = e.Kf * c1;
                e.thr = r.thr;
                e.lo = lo;
                e.hi = hi;
                e.nu = r.nu;
                e.dw = r.dw;
                e.y = r.y;
                e.K = r.K;
                // the hull of the level-0 tiles to evaluate: does the window cover it, does it lie entirely in region I?
                const int i0 = __ffs(msk) - 1, i1 = 31 - __clz(msk);
                const int64_t h0p = ws + (int64_t)i0 * SUB;
                const int64_t h1p = (ws + (int64_t)(i1 + 1) * SUB < we) ? ws + (int64_t)(i1 + 1) * SUB : we;
                if (lo <= h0p && hi >= h1p) {
                    double xa = fma(s_nsub[warp][2 * i0], e.inv_dw, -e.xl);
                    double xb = fma(s_nsub[warp][2 * i1 + 1], e.inv_dw, -e.xl);
                    ff = (xa * xb > 0.0) && (fmin(xa * xa, xb * xb) > e.thr);
                }
            }
            const unsigned m_far = __ballot_sync(0xffffffffu, pass && ff);
            const unsigned m_mix = __ballot_sync(0xffffffffu, pass && !ff);
            const int n_far = __popc(m_far), n_mix = __popc(m_mix);
            if (pass) {
                const int slot = ff ? __popc(m_far & lt_mask) : 31 - __popc(m_mix & lt_mask);
                my[slot] = e;
                my_msk[slot] = (unsigned char)msk;
            }
            __syncwarp();
            // ---- consume the far-wing entries ---------------------------------------------------------
            for (int k = 0; k < n_far; k++) {
                const WEntry &w = my[k];
                const unsigned mk = my_msk[k];
                far_eval(w.xl, w.inv_dw, w.b, w.c, w.Kc, w.Kf, mk);
                if (STATS) {
#pragma unroll
                    for (int i = 0; i < NSUB; i++)
                        if ((mk >> i) & 1u) h0 += (nvalid_sub >> (8 * i)) & 0xffu;
                }
            }
            // ---- mixed entries: window edge inside the span and/or pixels near the line core.  Lanes take CONSECUTIVE
            // pixels of the in-window part of the span (a 20-pixel window keeps 20 lanes busy in one pass); entries are
            // processed one after the other and a pass touches distinct pixels, so the shared accumulators need no
            // atomics and the summation order stays fixed.
            for (int m = 0; m < n_mix; m++) {
                const WEntry &e2 = my[31 - m];
                const unsigned mk = my_msk[31 - m];
                const int64_t pa = e2.lo > ws ? e2.lo : ws, pb = e2.hi < we ? e2.hi : we;
                const double thr = e2.thr, xl = e2.xl, inv_dw = e2.inv_dw, eb = e2.b, ec = e2.c, Kc = e2.Kc, Kf = e2.Kf;
                if (pb - pa <= 64) {
                    for (int64_t c0 = pa; c0 < pb; c0 += 32) {
                        const int64_t pix = c0 + lane;
                        if (pix < pb && ((mk >> ((int)(pix - ws) >> SD_FAR_TILE0_SHIFT)) & 1u)) {
                            const int k = (int)(pix - ws);
                            const double nu = nus[pix];
                            double x = fma(nu, inv_dw, -xl);
                            double q = x * x;
                            double v;
                            if (q > thr) {
                                double den = fma(q, q + eb, ec);
                                double num = fma(Kf, q, Kc);
                                v = num * (RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den));
                            } else {
                                v = exact_contribution(nu, e2.nu, e2.dw, inv_dw, thr, e2.y, e2.K);
                            }
                            s_acc[warp][k] += v;
                            if (STATS && pix >= p0 && pix < p1) {
                                int r = sdm::humlicek_region((nu - e2.nu) / e2.dw, e2.y);
                                h0 += (r == 0); h1 += (r == 1); h2 += (r == 2); h3 += (r == 3);
                            }
                        }
                    }
                    // the next entry's pass may update the same pixel from another lane: order the shared read-modify-
                    // writes of the warp (independent thread scheduling gives no lock-step guarantee after the divergent
                    // exact path; compute-sanitizer racecheck flagged exactly this line)
                    __syncwarp();
                } else if (e2.lo <= ws && e2.hi >= we && we - ws == SPAN) {
                    // long overlap, window covers the whole span (a near-field pair whose core lies in this span, the
                    // common case): register slots without any window logic
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        if (!((mk >> (p / SLOTS_PER_SUB)) & 1u)) continue;  // warp-uniform
                        const double x = fma(nu_i[p], inv_dw, -xl);
                        const double q = x * x;
                        const bool fast = q > thr;
                        const double den = fma(q, q + eb, ec);
                        const double num = fma(Kf, q, Kc);
                        const double v = num * (RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den));
                        if (fast) acc[p] += v;
                        if (!__all_sync(0xffffffffu, fast)) {
                            if (!fast) acc[p] += exact_contribution(nu_i[p], e2.nu, e2.dw, inv_dw, thr, e2.y, e2.K);
                        }
                        if (STATS) {
                            const int64_t pix = ws + p * 32 + lane;
                            if (pix >= p0 && pix < p1) {
                                int r = sdm::humlicek_region((nu_i[p] - e2.nu) / e2.dw, e2.y);
                                h0 += (r == 0); h1 += (r == 1); h2 += (r == 2); h3 += (r == 3);
                            }
                        }
                    }
                } else {
                    // long overlap with a window edge inside the span: register slots with the window test
                    const int lo2 = e2.lo, hi2 = e2.hi;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        if (!((mk >> (p / SLOTS_PER_SUB)) & 1u)) continue;  // warp-uniform
                        int64_t pix = ws + p * 32 + lane;
                        bool inwin = (pix >= lo2) && (pix < hi2) && (pix < t1);
                        if (!__any_sync(0xffffffffu, inwin)) continue;
                        double x = fma(nu_i[p], inv_dw, -xl);
                        double q = x * x;
                        bool fast = inwin && (q > thr);
                        double den = fma(q, q + eb, ec);
                        double num = fma(Kf, q, Kc);
                        double v = num * (RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den));
                        if (fast) acc[p] += v;
                        if (inwin && !fast) acc[p] += exact_contribution(nu_i[p], e2.nu, e2.dw, inv_dw, thr, e2.y, e2.K);
                        if (STATS && inwin && pix >= p0 && pix < p1) {
                            int r = sdm::humlicek_region((nu_i[p] - e2.nu) / e2.dw, e2.y);
                            h0 += (r == 0); h1 += (r == 1); h2 += (r == 2); h3 += (r == 3);
                        }
                    }
                }
            }
            __syncwarp();
            }
        }
    }

    __syncwarp();
#pragma unroll
    for (int p = 0; p < P; p++) acc[p] += s_acc[warp][p * 32 + lane];

    // ---- far field: one polynomial per hierarchy level (Horner in t = (nu - nu_c) / h of that level's tile); at level 0
    // every pair of register slots has its own tile
    if constexpr (FARCAP) if (use_far && warp_has_pixels) {
        for (int lev = a.n_act - 1; lev >= 1; lev--) {
            const int tk = (int)(ws >> a.fg.tile_shift[lev]);
            const double nu_c = a.fg.geom[lev][3 * tk], h_k = a.fg.geom[lev][3 * tk + 1];
            const double inv_h = (h_k > 0.0) ? 1.0 / h_k : 0.0;  // a one-pixel tile has h = 0: only the constant term counts
            const double *__restrict__ C = a.far_coef[lev] + ((size_t)d * a.far_ntl[lev] + (tk - a.far_tile0[lev])) * K1;
            double tt[P], poly[P];
            const double ck = C[K1 - 1];
#pragma unroll
            for (int p = 0; p < P; p++) {
                tt[p] = (nu_i[p] - nu_c) * inv_h;
                poly[p] = ck;
            }
#pragma unroll
            for (int k = K1 - 2; k >= 0; k--) {
                const double c = C[k];
#pragma unroll
                for (int p = 0; p < P; p++) poly[p] = fma(poly[p], tt[p], c);
            }
#pragma unroll
            for (int p = 0; p < P; p++) acc[p] += poly[p];
        }
#pragma unroll
        for (int i = 0; i < NSUB; i++) {
            const int64_t c0 = ws + (int64_t)i * SUB;
            if (c0 >= we || c0 + SUB <= p0 || c0 >= p1) continue;  // warp-uniform; tiles outside the shard have no coefficients
            const int tk = (int)(c0 >> SD_FAR_TILE0_SHIFT);
            const double nu_c = a.fg.geom[0][3 * tk], h_k = a.fg.geom[0][3 * tk + 1];
            const double inv_h = (h_k > 0.0) ? 1.0 / h_k : 0.0;
            const double *__restrict__ C = a.far_coef[0] + ((size_t)d * a.far_ntl[0] + (tk - a.far_tile0[0])) * K1;
            double tt[SLOTS_PER_SUB], poly[SLOTS_PER_SUB];
            const double ck = C[K1 - 1];
#pragma unroll
            for (int q = 0; q < SLOTS_PER_SUB; q++) {
                tt[q] = (nu_i[i * SLOTS_PER_SUB + q] - nu_c) * inv_h;
                poly[q] = ck;
            }
#pragma unroll
            for (int k = K1 - 2; k >= 0; k--) {
                const double c = C[k];
#pragma unroll
                for (int q = 0; q < SLOTS_PER_SUB; q++) poly[q] = fma(poly[q], tt[q], c);
            }
#pragma unroll
            for (int q = 0; q < SLOTS_PER_SUB; q++) acc[i * SLOTS_PER_SUB + q] += poly[q];
        }
    }

#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        if (pix < t1 && pix >= p0 && pix < p1) a.out[(size_t)d * (p1 - p0) + (pix - p0)] = acc[p];
    }
    if (STATS) {
        for (int o2 = 16; o2; o2 >>= 1) {
            h0 += __shfl_xor_sync(0xffffffffu, h0, o2);
            h1 += __shfl_xor_sync(0xffffffffu, h1, o2);
            h2 += __shfl_xor_sync(0xffffffffu, h2, o2);
            h3 += __shfl_xor_sync(0xffffffffu, h3, o2);
        }
        if (lane == 0) {
            if (h0) atomicAdd(&a.stats[0], h0);
            if (h1) atomicAdd(&a.stats[1], h1);
            if (h2) atomicAdd(&a.stats[2], h2);
            if (h3) atomicAdd(&a.stats[3], h3);
        }
    }
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int P, int NW>
int launch(sd_ctx *c, const LineArgs &a, dim3 grid, bool stats, int rcp) {
    constexpr int MINB = (NW == 8) ? 2 : (NW == 2 ? 8 : 16);  // <= 128 registers per thread in every configuration
    // the counting instantiation uses the production arithmetic (Newton reciprocal) so that both are bitwise equal
    if (stats) k_lines<P, NW, true, 2, MINB><<<grid, 32 * NW, 0, c->stream>>>(a);
    else if (rcp == 3) k_lines<P, NW, false, 3, MINB><<<grid, 32 * NW, 0, c->stream>>>(a);
    else {
        static const int minb = env_int("SD_K2_MINB", 0);  // tuning experiments: more resident CTAs, fewer registers
        if (NW == 2 && P == 8 && minb == 10) k_lines<8, 2, false, 2, 10><<<grid, 64, 0, c->stream>>>(a);
        else if (NW == 2 && P == 8 && minb == 12) k_lines<8, 2, false, 2, 12><<<grid, 64, 0, c->stream>>>(a);
        else k_lines<P, NW, false, 2, MINB><<<grid, 32 * NW, 0, c->stream>>>(a);
    }
    return sd_launch_check(c, "k_lines");
}

}  // namespace

// pixels per thread: enough CTAs to fill the chip several times over, otherwise as much register reuse of the staged
// entries as possible.  The choice is made from the GLOBAL grid length, never from the shard: the tile size fixes the
// summation order inside a pixel, and a nu shard must reproduce the columns of the full run bit for bit.
// SD_K2_P / SD_K2_NW override the choice (tuning experiments).
int sd_k2_choose_P(sd_ctx *c) {
    static const int force_p = env_int("SD_K2_P", 0);
    static const int force_nw = env_int("SD_K2_NW", 0);
    const int64_t W = c->N;
    int P, NW = 8;
    if (c->farfield) {
        // The hierarchy absorbs everything but the level-0 tiles around a line; wide per-warp spans keep the per-pair
        // staging amortised: 2 warps x 256 pixels (1 warp on small grids).
        P = 8;
        NW = (((W + 511) / 512) * c->D >= 2LL * c->sm_count) ? 2 : 1;
    } else if (((W + 2047) / 2048) * c->D >= 8LL * c->sm_count) P = 8;
    else if (((W + 1023) / 1024) * c->D >= 4LL * c->sm_count) P = 4;
    else if (((W + 511) / 512) * c->D >= 4LL * c->sm_count) P = 2;
    else P = 1;
    if (force_p == 1 || force_p == 2 || force_p == 4 || force_p == 8) P = force_p;
    if (force_nw == 1 || force_nw == 2 || force_nw == 8) NW = force_nw;
    if (c->farfield) {  // level-0 tiles are pairs of register slots and a warp span must not straddle a level-1 tile
        P = 8;
        if (NW == 8) NW = 2;
    }
    if (NW != 8 && P != 8) P = 8;  // only P = 8 is instantiated for the narrow CTAs
    c->k2_NW = NW;
    return P;
}

int sd_k2_lines(sd_ctx *c, int slot) {
    const int64_t W = c->W();
    SD_TRY(sd_ensure(c, c->alpha_line[slot], sizeof(double) * c->D * W));
    if (c->L == 0) {
        SD_CUDA(c, cudaMemsetAsync(c->alpha_line[slot].p, 0, sizeof(double) * c->D * W, c->stream));
        return SD_OK;
    }
    if (c->line_stats) {
        SD_CUDA(c, cudaMemsetAsync(c->stats.p, 0, 4 * sizeof(unsigned long long), c->stream));
        SD_CUDA(c, cudaMemsetAsync(c->stats.as<unsigned long long>() + 8, 0, 3 * sizeof(unsigned long long), c->stream));
        SD_CUDA(c, cudaMemsetAsync(c->stats.as<unsigned long long>() + 12, 0, 2 * sizeof(unsigned long long), c->stream));
    }
    static const int rcp = env_int("SD_K2_RCP", 2);
    const int P = c->k2_P, NW = c->k2_NW, tile = 32 * NW * P;
    LineArgs a{};
    a.L = c->L; a.N = c->N; a.p0 = c->p0; a.p1 = c->p1; a.D = c->D;
    a.tile0 = (int)(c->p0 / tile);
    a.n_tiles = (int)((c->N + tile - 1) / tile);
    const int n_launch = (int)((c->p1 + tile - 1) / tile) - a.tile0;
    a.nus = c->nus.as<double>(); a.line_idx = c->line_idx.as<int>(); a.rec = c->rec.as<LineRec>();
    a.win = c->win.as<PairWin>();
    a.cls_list = c->cls_list.as<int>(); a.cls_off = c->cls_off.as<int>();
    a.fg = c->far_geom;
    a.n_act = c->farfield ? c->far_active : 0;
    a.out = c->alpha_line[slot].as<double>();
    a.stats = c->stats.as<unsigned long long>();
    if (c->farfield && a.n_act > 0) {
        const int n_act = a.n_act;
        for (int k = n_act - 1; k >= 0; k--) {
            const int tk = a.fg.tile[k];
            a.far_tile0[k] = (int)(c->p0 / tk);
            a.far_ntl[k] = (int)((c->p1 + tk - 1) / tk) - a.far_tile0[k];
            SD_TRY(sd_ensure(c, c->far_coef[k], sizeof(double) * c->D * a.far_ntl[k] * K1));
            a.far_coef[k] = c->far_coef[k].as<double>();
            // multipole moments: every tile of every level (the top level reaches the whole grid)
            SD_TRY(sd_ensure(c, c->far_mom[k], sizeof(double) * c->D * a.fg.n_tiles[k] * K1));
            a.far_mom[k] = c->far_mom[k].as<double>();
            for (int h = k; h < n_act; h++) {
                SD_TRY(sd_ensure(c, c->far_bkt[k][h], sizeof(double) * c->D * a.fg.n_tiles[k] * K1));
                a.far_bkt[k][h] = c->far_bkt[k][h].as<double>();
            }
        }
        if (!c->far_attr_set) {  // per device: > 48 KB of dynamic shared memory needs the opt-in; series-length table
            SD_CUDA(c, cudaFuncSetAttribute(k_far_coeffs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAR_SMEM));
            unsigned char tab[256];
            far_terms_table(tab);
            SD_CUDA(c, cudaMemcpyToSymbolAsync(FAR_TERMS, tab, sizeof tab, 0, cudaMemcpyHostToDevice, c->stream));
            double inv[K1 + 2];
            inv[0] = 0.0;
            for (int j = 1; j < K1 + 2; j++) inv[j] = 1.0 / (double)j;
            SD_CUDA(c, cudaMemcpyToSymbolAsync(M2M_INV, inv, sizeof inv, 0, cudaMemcpyHostToDevice, c->stream));
            SD_CUDA(c, cudaStreamSynchronize(c->stream));  // `tab` and `inv` live on this stack frame
            c->far_attr_set = true;
        }
        // CTAs per (group of eight sibling tiles, depth) of the direct expansion: the candidate lists are cut into this
        // many fixed slices so that even a narrow shard fills the chip; k_far_reduce adds the partial sums in slice order.
        size_t part_bytes = 0;
        for (int k = 0; k < n_act; k++) {
            const int ns = far_nsplit(k, k == n_act - 1);
            const size_t bb = ns > 1 ? sizeof(double) * c->D * a.far_ntl[k] * ns * K1 : 0;
            part_bytes = bb > part_bytes ? bb : part_bytes;
        }
        SD_TRY(sd_ensure(c, c->far_part, part_bytes > 0 ? part_bytes : 8));
        sd_phase_begin(c, SD_PH_FAR);
        const int cs = c->line_stats ? 1 : 0;
        // multipole moments: one expansion per saturated pair at its level lmin, then upwards tile to tile
        for (int m = 0; m < n_act; m++) {
            const int runs = (a.fg.n_tiles[m] + (1 << SD_FAR_SHIFT) - 1) >> SD_FAR_SHIFT;
            k_s2m<<<dim3((unsigned)((runs + S2M_WARPS - 1) / S2M_WARPS), (unsigned)c->D), 32 * S2M_WARPS, 0, c->stream>>>(a, m, cs);
            SD_TRY(sd_launch_check(c, "k_s2m"));
        }
        for (int lev = 1; lev <= n_act; lev++) {
            const int parents = (a.fg.n_tiles[lev - 1] + (1 << SD_FAR_SHIFT) - 1) >> SD_FAR_SHIFT;
            k_m2m<<<dim3((unsigned)((parents + 3) / 4), (unsigned)c->D), 128, 0, c->stream>>>(a, lev);
            SD_TRY(sd_launch_check(c, "k_m2m"));
        }
        for (int k = n_act - 1; k >= 0; k--) {
            const int nsplit = far_nsplit(k, k == n_act - 1);
            // direct expansions: one CTA per group of eight sibling tiles that has a member in the launched range (times
            // the slices of the candidate lists; k_far_reduce adds the partial sums in slice order)
            const int n_grp = ((a.far_tile0[k] + a.far_ntl[k] - 1) >> SD_FAR_SHIFT) - (a.far_tile0[k] >> SD_FAR_SHIFT) + 1;
            k_far_coeffs<<<dim3((unsigned)(n_grp * nsplit), (unsigned)c->D), THREADS, FAR_SMEM, c->stream>>>(
                a, k, cs, nsplit, c->far_part.as<double>());
            SD_TRY(sd_launch_check(c, "k_far_coeffs"));
            if (nsplit > 1) {
                const int n = c->D * a.far_ntl[k] * K1;
                k_far_reduce<<<(n + 255) / 256, 256, 0, c->stream>>>(n, nsplit, c->far_part.as<double>(), a.far_coef[k]);
                SD_TRY(sd_launch_check(c, "k_far_reduce"));
            }
            // depth chunks of 14, or of 8 when that wastes fewer accumulator slots (a depth-sharded rank holds D / R depths)
            const int w14 = ((c->D + 13) / 14) * 14 - c->D, w8 = ((c->D + 7) / 8) * 8 - c->D;
            if (w8 < w14) k_m2l<8><<<dim3((unsigned)n_grp, (unsigned)((c->D + 7) / 8)), THREADS, 0, c->stream>>>(a, k, cs);
            else k_m2l<M2L_DC_MAX><<<dim3((unsigned)n_grp, (unsigned)((c->D + 13) / 14)), THREADS, 0, c->stream>>>(a, k, cs);
            SD_TRY(sd_launch_check(c, "k_m2l"));
        }
        sd_phase_end(c, SD_PH_FAR);
    }
    dim3 grid((unsigned)n_launch, (unsigned)c->D);
    sd_phase_begin(c, SD_PH_LINES);
    int rc;
    if (NW == 2) rc = launch<8, 2>(c, a, grid, c->line_stats, rcp);
    else if (NW == 1) rc = launch<8, 1>(c, a, grid, c->line_stats, rcp);
    else switch (P) {
        case 8: rc = launch<8, 8>(c, a, grid, c->line_stats, rcp); break;
        case 4: rc = launch<4, 8>(c, a, grid, c->line_stats, rcp); break;
        case 2: rc = launch<2, 8>(c, a, grid, c->line_stats, rcp); break;
        default: rc = launch<1, 8>(c, a, grid, c->line_stats, rcp); break;
    }
    sd_phase_end(c, SD_PH_LINES);
    return rc;
}
