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
//     32-ary warp binary searches on the monotone window centres);
//   * k_lines: every WARP streams the candidates in batches of 32 on its own (no CTA barrier in the main loop): test
//     against the warp's 32*P-pixel span, expand the passing (line, depth) records into 96-byte shared-memory entries
//     (constants hoisted once per warp), entries whose window covers the span and whose span lies entirely in
//     Humlicek region I packed first ("far-wing" list), the others from the back ("mixed" list), then consume them
//     with broadcast LDS; summation order per pixel is fixed;
//   * far-wing hot loop (region I):  Kf (q + c1) / (q (q + b) + c),  q = x^2:  8 FP64 instructions + 1 MUFU.RCP64H
//     per evaluation (x, q, 2 for the denominator, numerator, 2 for the Newton step on the reciprocal seed,
//     accumulate), no branches, no divisions;
//   * pixels that are not certainly in region I take the exact path: x = dnu / dw (IEEE division) and the
//     reference's own region tests, so the Humlicek classification is identical to the reference's;
//   * FAR FIELD (k_far_coeffs + polynomial epilogue of k_lines): a pair whose window covers the whole tile and whose
//     line centre is at least 4 tile half-widths away contributes a function that is analytic over the tile.  In
//     region I,  Re w = (1/(2 sqrt(pi))) [ y/((x-a)^2+y^2) + y/((x+a)^2+y^2) ],  a = 1/sqrt(2): two Lorentzians, i.e.
//     the imaginary part of two simple poles p = nu_l -+ dw/sqrt(2) + i y dw.  Their Taylor series about the tile
//     centre nu_c,  1/(nu - p) = sum_k (-1)^k (nu - nu_c)^k / (nu_c - p)^(k+1),  converges with ratio <= 1/4; degree
//     20 reproduces the direct evaluation to <= 6e-12 (relative, worst case; all terms are positive).
//     k_far_coeffs accumulates the 21 coefficients of ALL far pairs of a tile (one pair per thread, ~230 FP64
//     operations instead of 8 per pixel), k_lines skips exactly those pairs (same integer test on the per-pair
//     "near tile interval" computed by k_build_records) and adds the polynomial at the end.
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
struct __align__(16) FarRec { double nu, dw, y, K; };  // = the first 32 bytes of LineRec
static_assert((SD_FAR_K + 1) % 3 == 0, "the series length is tested every third term");
static_assert(WARPS == (1 << SD_FAR_SHIFT), "k_far_coeffs maps the warps of a CTA to the children of a tile");
constexpr size_t FAR_SMEM = (size_t)FAR_CH * sizeof(FarRec) + FAR_CH;

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
    int tile0;     // global index of the first tile of this launch
    int n_tiles;   // global number of tiles
    const double *nus;
    const int *line_idx;
    const LineRec *rec;
    const PairWin *win;   // (depth, line) window records
    const int *cls_list, *cls_off;
    FarGeom fg;                  // tile hierarchy; fg.near[0] == nullptr: far field disabled
    double *far_coef[SD_FAR_LEVELS];   // per level: (D, n_tiles_launch[k], SD_FAR_K + 1)
    int far_tile0[SD_FAR_LEVELS];      // first global tile of this launch, per level
    int far_ntl[SD_FAR_LEVELS];        // tiles of this launch, per level
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
    if (cls >= SD_NCLS - 2) {  // classes 6 (unbounded half-width) and 7 (whole grid): every pair is a candidate
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

// first j in [a, b) with keys[j] >= X (keys ascending), b if none.  Warp-cooperative 32-ary search.
__device__ __forceinline__ int warp_lower_bound_u64(const unsigned long long *__restrict__ keys, int a, int b, unsigned long long X) {
    const int lane = threadIdx.x & 31;
    while (b > a) {
        int n = b - a;
        int step = (n + 31) >> 5;
        long long pj = (long long)a + (long long)lane * step;
        bool pred = (pj >= b) ? true : (keys[pj] >= X);
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

// Far-capable pairs (class 7, sorted by window centre) whose centre lies within `rad` level-`lev` tiles of tile `t`.
__device__ __forceinline__ void fc_near_range(const LineArgs &a, int d, int lev, int t, int &ja, int &jb) {
    const int lo = a.cls_off[d * (SD_NCLS + 1) + SD_FC_CLASS], hi = a.cls_off[d * (SD_NCLS + 1) + SD_FC_CLASS + 1];
    const int *list_d = a.cls_list + (size_t)d * a.L;
    const int *line_idx = a.line_idx;
    const long long T = a.fg.tile[lev], rad = a.fg.near_rad[lev];
    auto key = [&](int j) { return line_idx[list_d[j]]; };
    ja = warp_first_below(key, lo, hi, clamp_i32(((long long)t + rad + 1) * T));  // centre <  (t + rad + 1) T
    jb = warp_first_below(key, ja, hi, clamp_i32(((long long)t - rad) * T));      // centre <  (t - rad) T
}

// Far-capable pairs of depth d with a window start (which = 0) or end (which = 1) strictly inside (t0, t1).
__device__ __forceinline__ void fc_edge_range(const LineArgs &a, int d, int which, int64_t t0, int64_t t1, int &ja, int &jb) {
    const int lo = a.fg.edge_off[which * (a.D + 1) + d], hi = a.fg.edge_off[which * (a.D + 1) + d + 1];
    ja = warp_lower_bound_u64(a.fg.edge_keys, lo, hi, sd_edge_key(a.fg, which, d, t0 + 1, 0));
    jb = warp_lower_bound_u64(a.fg.edge_keys, ja, hi, sd_edge_key(a.fg, which, d, t1, 0));
}

// one 32-byte gather (two 16-byte loads of the same sector)
__device__ __forceinline__ PairWin load_win(const PairWin *__restrict__ w) {
    const int4 a = __ldg(reinterpret_cast<const int4 *>(w)), b = __ldg(reinterpret_cast<const int4 *>(w) + 1);
    PairWin r;
    r.lo = a.x; r.hi = a.y; r.near[0] = (unsigned)a.z; r.near[1] = (unsigned)a.w;
    r.near[2] = (unsigned)b.x; r.cls = b.y; r.pad0 = 0; r.pad1 = 0;
    return r;
}

__device__ __forceinline__ unsigned near_of(const PairWin &w, int lev) {  // no dynamically indexed registers
    return lev == 0 ? w.near[0] : (lev == 1 ? w.near[1] : w.near[2]);
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

// The same integer test in both kernels: the pair covers the whole global tile and the tile lies outside the pair's
// near interval -> it is expanded (k_far_coeffs) and must be skipped by the direct kernel.
__device__ __forceinline__ bool pair_is_far(int lo, int hi, unsigned near, int64_t t0, int64_t t1, int tile) {
    const int nl = (int)(near & 0xffffu), nh = (int)(near >> 16);
    return (lo <= t0) && (hi >= t1) && (tile < nl || tile >= nh);
}

// ------------------------------------------------------------------------------------------------------------------
// Far-field coefficients of one (level-`lev` tile, depth): C_k = -W Im(w+^(k+1) + w-^(k+1)),  w = -h / (nu_c - p),
// W = K dw / (2 sqrt(pi) h);  the contribution of the pair at pixel nu is  sum_k C_k ((nu - nu_c)/h)^k.
// A pair is expanded at the HIGHEST level at which it is far: level `lev` takes the far-capable pairs that cover this
// tile, are far from it, and are NOT far for the parent tile of level lev + 1.  Those are found without scanning:
//   (A) pairs that cover the parent but have it in their near interval: a contiguous range (by window centre) of the
//       class-7 list around the parent;
//   (B) pairs that do not cover the parent (a window edge lies strictly inside it): two ranges of the edge-sorted lists.
// The candidates are a property of the PARENT, so one CTA serves the eight children of a parent (warp w = child w):
//   scan     the CTA walks the candidate lists in chunks of FAR_CH; a thread gathers the 32-byte window record of a
//            candidate ONCE, tests it against all eight children (integer compares) and leaves an 8-bit acceptance mask;
//            the 32 bytes of LineRec the expansion needs are staged in shared memory with cp.async if any child wants
//            them (one gather per candidate and parent instead of one per candidate and child);
//   expand   every warp picks its child's bit out of the masks, compacts the accepted records into its own queue in list
//            order (ballots) and expands full batches of 32, one pair per lane, all lanes busy.
// The top level has no parent: groups of eight consecutive tiles walk a fixed slice of the whole class-7 list
// (`nsplit` CTAs per group; k_far_reduce adds the partial sums in slice order).  Groups, slices, chunking and queue order
// depend on the global tile index and the candidate lists only, never on the shard, so the summation order -- and the
// result, bit for bit -- is the same for every partition of the grid.
// series length by floor(-8 log2(rho^2)) (see `expand`): filled by far_terms_table(), uploaded once per device
__constant__ unsigned char FAR_TERMS[256];

void far_terms_table(unsigned char *tab) {
    constexpr int K1 = SD_FAR_K + 1;
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

__global__ void __launch_bounds__(THREADS) k_far_coeffs(LineArgs a, int lev, int count_stats, int nsplit, double *part) {
    constexpr int K1 = SD_FAR_K + 1;
    __shared__ int s_ja[3], s_jb[3];
    extern __shared__ __align__(16) unsigned char far_smem[];
    FarRec *const s_rec = reinterpret_cast<FarRec *>(far_smem);                        // [FAR_CH] dense records of the chunk
    unsigned char *const s_mask = reinterpret_cast<unsigned char *>(s_rec + FAR_CH);   // [FAR_CH] their child masks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = blockIdx.y;
    const int tile_px = a.fg.tile[lev];
    const bool has_parent = lev + 1 < SD_FAR_LEVELS;
    const int plev = has_parent ? lev + 1 : lev;
    const int group = (a.far_tile0[lev] >> SD_FAR_SHIFT) + (int)blockIdx.x / nsplit;  // = parent tile index
    const int split = (int)blockIdx.x % nsplit;
    const int child0 = group << SD_FAR_SHIFT;
    const int tile_w = child0 + warp;
    const bool tile_ok = tile_w >= a.far_tile0[lev] && tile_w < a.far_tile0[lev] + a.far_ntl[lev];
    const int tile = tile_ok ? tile_w : a.far_tile0[lev];
    const int64_t t0 = (int64_t)tile * tile_px;
    const int64_t t1 = (t0 + tile_px < a.N) ? t0 + tile_px : a.N;
    const double nu_c = a.fg.geom[lev][2 * tile], h = a.fg.geom[lev][2 * tile + 1];
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
    if (warp < 3) {
        int ja, jb;
        if (!has_parent) {
            ja = a.cls_off[d * (SD_NCLS + 1) + SD_FC_CLASS];
            jb = (warp == 0) ? a.cls_off[d * (SD_NCLS + 1) + SD_FC_CLASS + 1] : ja;
        } else if (warp == 0) {
            fc_near_range(a, d, plev, ptile, ja, jb);
        } else {
            fc_edge_range(a, d, warp - 1, pt0, pt1, ja, jb);
        }
        // this CTA's slice of the range (a function of the range and nsplit only)
        const int len = (jb - ja + nsplit - 1) / nsplit;
        const int sa = min(ja + split * len, jb), sb = min(sa + len, jb);
        if (lane == 0) { s_ja[warp] = sa; s_jb[warp] = sb; }
    }
    __syncthreads();
    double C[K1];
#pragma unroll
    for (int k = 0; k < K1; k++) C[k] = 0.0;
    unsigned long long n_far = 0, n_terms = 0;
    const double inv_h = 1.0 / h;
    const unsigned lt_mask = (1u << lane) - 1u;

    // One accepted pair: 21 Taylor coefficients of its two poles about the tile centre.  Called with a dense batch of
    // pairs (one per lane); `have` is false only in the last, partial batch.
    auto expand = [&](bool have, const FarRec &r) {
        double Wn = 0.0, w1r = 0.0, w1i = 0.0, w2r = 0.0, w2i = 0.0;
        int nterms = 0;
        if (have) {
            const double g = r.y * r.dw;                                // Lorentz half-width in Hz
            Wn = -r.K * r.dw * (0.5 * sdm::INV_SQRT_PI) * inv_h;        // -W
            const double adw = 0.7071067811865476 * r.dw;
            // w = -h / (D - i g) = -h (D + i g) / (D^2 + g^2) for the two poles
            const double D1 = nu_c - (r.nu + adw), D2 = nu_c - (r.nu - adw);
            const double q1 = sdm::rcp_fast(fma(D1, D1, g * g)), q2 = sdm::rcp_fast(fma(D2, D2, g * g));
            const double i1 = -h * q1, i2 = -h * q2;
            w1r = D1 * i1; w1i = g * i1; w2r = D2 * i2; w2i = g * i2;
            // terms needed: (n + 1) rho^n <= (K1 + 1) rho_far^K1 (the bound of the full series at the far criterion)
            // <=>  n >= ~K1 log2(1 / rho_far) / log2(1 / rho);  rho^2 = h^2 max(q1, q2).  -log2(rho^2) is read off the
            // exponent and the top mantissa bits of rho^2 in steps of 1/8 (a lower bound: the series is never shorter
            // than the rule asks) and indexes a 256-entry table -- six integer instructions instead of ~30 with the
            // float logarithm and division this used to be (9 % of the kernel's instructions, ncu round 2)
            const int t8 = (0x3ff00000 - __double2hiint(h * h * fmax(q1, q2))) >> 17;  // floor(8 * -L), L <= log2(rho^2) <= L + 0.086
            nterms = FAR_TERMS[min(max(t8, 0), 255)];
            if (count_stats) n_far++;
        }
        // queue neighbours are neighbours in frequency, at similar distances from the tile: warp-uniform series length
        const int nt = __reduce_max_sync(0xffffffffu, nterms);
        if (count_stats && have) n_terms += (unsigned long long)min(K1, 3 * ((nt + 2) / 3));  // terms the loop below executes
        // Im(w^(k+1)) by the real three-term recurrence of the powers of a complex number,
        //   s_(k+1) = 2 Re(w) s_k - |w|^2 s_(k-1),  s_0 = 0, s_1 = Im w,
        // two instructions per pole and term instead of the four of a complex product (the recurrence loses about one
        // bit per step relative to |w|^k, i.e. < 1e-13 over 21 terms).
        const double a1 = w1r + w1r, b1 = fma(w1r, w1r, w1i * w1i), a2 = w2r + w2r, b2 = fma(w2r, w2r, w2i * w2i);
        double s1 = w1i, s1p = 0.0, s2 = w2i, s2p = 0.0;
#pragma unroll
        for (int k = 0; k < K1; k++) {
            if (k % 3 == 0 && k >= nt) break;  // checked every third term (the extra terms only add accuracy)
            C[k] = fma(Wn, s1 + s2, C[k]);
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
    //            the shard) are packed densely, in list order, into shared memory: record (cp.async) + child mask.  Far
    //            from the group every covering pair is wanted by all eight children, near it by five or six, so the
    //            dense list is (nearly) the list of every child;
    //   expand   every warp walks the dense list 32 entries at a time, one pair per lane, skipping the few entries whose
    //            mask lacks its child -- no per-warp queue, no ballots, no copies (they were a quarter of the kernel's
    //            instructions when every warp compacted its own list).
    constexpr int ROUNDS = FAR_CH / THREADS;
    constexpr int SEGS = ROUNDS * WARPS;       // (round, warp) segments of 32 candidates, in list order
    static_assert(SEGS == 32, "one lane per segment in the offset scan");
    __shared__ int s_cnt[SEGS];
    for (int src = 0; src < 3; src++) {
        const int ja = s_ja[src], jb = s_jb[src];
        for (int base = ja; base < jb; base += FAR_CH) {
            // ---- test
            unsigned mk[ROUNDS];
            int ll[ROUNDS], rk[ROUNDS];
#pragma unroll
            for (int r = 0; r < ROUNDS; r++) {
                const int idx = r * THREADS + tid, j = base + idx;
                unsigned mask = 0;
                int l = 0;
                if (j < jb) {
                    l = (src == 0) ? list_d[j] : (int)(a.fg.edge_keys[j] & lmask);
                    const PairWin pw = load_win(a.win + drow + l);
                    const int lo = pw.lo, hi = pw.hi;
                    bool okp = true;
                    if (has_parent) {
                        const bool covers_parent = (lo <= pt0) && (hi >= pt1);
                        if (src == 0) {  // (A): covers the parent, parent inside the near interval
                            okp = covers_parent && !pair_is_far(lo, hi, near_of(pw, plev), pt0, pt1, ptile);
                        } else {         // (B): an edge strictly inside the parent; both edges inside: via its start
                            okp = !covers_parent && !(src == 2 && lo > pt0 && lo < pt1);
                        }
                    }
                    if (okp) {
                        const unsigned near = near_of(pw, lev);
#pragma unroll
                        for (int cc = 0; cc < WARPS; cc++) {
                            const int tcc = child0 + cc;
                            const int64_t c0 = (int64_t)tcc * tile_px;
                            const int64_t c1 = (c0 + tile_px < a.N) ? c0 + tile_px : a.N;
                            if (tcc < a.fg.n_tiles[lev] && pair_is_far(lo, hi, near, c0, c1, tcc)) mask |= 1u << cc;
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
    // deterministic reduction: lanes by shuffle (every lane ends up with the sum; lane k keeps coefficient k)
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < K1; k++) {
        double v = C[k];
        for (int o2 = 16; o2; o2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o2);
        if (lane == k) mine = v;
    }
    if (tile_ok && lane < K1) {
        const size_t tl = (size_t)d * a.far_ntl[lev] + (tile - a.far_tile0[lev]);
        if (nsplit > 1) part[(tl * nsplit + split) * K1 + lane] = mine;
        else a.far_coef[lev][tl * K1 + lane] = mine;
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
__host__ __device__ constexpr int far_nsplit_base(int lev) { return lev == SD_FAR_LEVELS - 1 ? 32 : (lev == 0 ? 2 : 8); }
// SD_FAR_NSPLIT_SCALE (tuning experiments only: it changes the grouping of the partial sums, i.e. the last bits)
static int far_nsplit(int lev) {
    static const int scale = env_int_early("SD_FAR_NSPLIT_SCALE", 1);
    return far_nsplit_base(lev) * (scale >= 1 && scale <= 8 ? scale : 1);
}

// sum of the nsplit partial coefficient sets of a level, in slice order
__global__ void k_far_reduce(int n, int nsplit, const double *__restrict__ part, double *__restrict__ coef) {
    constexpr int K1 = SD_FAR_K + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // (depth, tile) * K1 + k
    if (i >= n) return;
    const int td = i / K1, k = i - td * K1;
    double v = 0.0;
    for (int s2 = 0; s2 < nsplit; s2++) v += part[((size_t)td * nsplit + s2) * K1 + k];
    coef[i] = v;
}

// ------------------------------------------------------------------------------------------------------------------
template <int P, int NW, bool STATS, int RCP, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) k_lines(LineArgs a) {
    constexpr int TILE = 32 * NW * P;
    constexpr int SPAN = 32 * P;
    __shared__ WEntry s_ent[NW][32];  // every warp streams its own batches: no CTA barrier in the main loop
    __shared__ double s_acc[NW][SPAN];  // accumulators of the pixel-parallel mixed path
    constexpr int NSRC = SD_NCLS + 2;    // classes 0..6, far-capable pairs near the tile, their window starts / ends
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
    const bool use_far = a.fg.enabled != 0;
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
    // frequencies at the two ends of this warp's span (x is monotone in the pixel index)
    const double nu_first = nus[ws < N ? ws : N - 1];
    const double nu_last = nus[(we - 1 >= ws && we - 1 < N) ? we - 1 : (ws < N ? ws : N - 1)];

    // candidate ranges of all sources, distributed over the warps of the CTA
    for (int src = warp; src < NSRC; src += NW) {
        int ja = 0, jb = 0;
        if (src < SD_FC_CLASS) class_range(a, d, src, t0, t1, ja, jb);
        else if (src == SD_FC_CLASS) {
            if (use_far) fc_near_range(a, d, 0, tile, ja, jb);
        } else if (use_far) fc_edge_range(a, d, src - SD_NCLS, t0, t1, ja, jb);
        if (lane == 0) { s_ja[src] = ja; s_jb[src] = jb; }
    }
    __syncthreads();

    unsigned long long h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    int nvalid = 0;  // this lane's pixels that belong to the shard (statistics only)
#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        nvalid += (pix < t1) && (pix >= p0) && (pix < p1);
    }

    WEntry *const my = s_ent[warp];
    const bool warp_has_pixels = (ws < t1) && (ws < p1) && (we > p0);

    // far-wing (region I) evaluation of one staged entry for the P pixels of this lane
    auto far_eval = [&](const double xl, const double inv_dw, const double eb, const double ec, const double Kc,
                        const double Kf) {
#pragma unroll
        for (int p = 0; p < P; p++) {
            double x = fma(nu_i[p], inv_dw, -xl);
            double q = x * x;
            double den = fma(q, q + eb, ec);
            double num = fma(Kf, q, Kc);
            acc[p] = fma(num, RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den), acc[p]);
        }
    };

    for (int src = 0; src < NSRC && warp_has_pixels; src++) {
        const int ja = s_ja[src], jb = s_jb[src];
        for (int base = ja; base < jb; base += 32) {
            {
            // ---- test the 32 candidates of this batch against THIS WARP's span ------------------------
            const int j = base + lane;
            bool pass = false;
            size_t o = 0;
            int lo = 0, hi = 0;
            if (j < jb) {
                int l;
                if (src == 0) l = j;                                   // class 0: the nu-sorted line list itself
                else if (src <= SD_FC_CLASS) l = list_d[j];            // class lists (7 = far-capable pairs near the tile)
                else l = (int)(a.fg.edge_keys[j] & lmask);             // far-capable pairs with an edge inside the tile
                o = drow + l;
                const PairWin pw = load_win(a.win + o);
                lo = pw.lo;
                hi = pw.hi;
                pass = (lo < we) && (hi > ws) && (hi > lo);
                if (src == 0) pass = pass && (pw.cls == 0);
                else if (src == SD_FC_CLASS) {
                    // covering pairs only (the others come through the edge lists); skip those expanded by k_far_coeffs
                    pass = pass && (lo <= t0) && (hi >= t1) && !pair_is_far(lo, hi, pw.near[0], t0, t1, tile);
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
                if (lo <= ws && hi >= we) {  // window covers the whole span: is the span entirely in region I?
                    double xa = fma(nu_first, e.inv_dw, -e.xl);
                    double xb = fma(nu_last, e.inv_dw, -e.xl);
                    ff = (xa * xb > 0.0) && (fmin(xa * xa, xb * xb) > e.thr);
                }
            }
            const unsigned m_far = __ballot_sync(0xffffffffu, pass && ff);
            const unsigned m_mix = __ballot_sync(0xffffffffu, pass && !ff);
            const int n_far = __popc(m_far), n_mix = __popc(m_mix);
            if (pass) my[ff ? __popc(m_far & lt_mask) : 31 - __popc(m_mix & lt_mask)] = e;
            __syncwarp();
            // ---- consume the far-wing entries ---------------------------------------------------------
            for (int k = 0; k < n_far; k++) {
                const WEntry &w = my[k];
                far_eval(w.xl, w.inv_dw, w.b, w.c, w.Kc, w.Kf);
            }
            if (STATS) h0 += (unsigned long long)n_far * nvalid;
            // ---- mixed entries: window edge inside the span and/or pixels near the line core.  Lanes take CONSECUTIVE
            // pixels of the in-window part of the span (a 20-pixel window keeps 20 lanes busy in one pass); entries are
            // processed one after the other and a pass touches distinct pixels, so the shared accumulators need no
            // atomics and the summation order stays fixed.
            for (int m = 0; m < n_mix; m++) {
                const WEntry &e2 = my[31 - m];
                const int64_t pa = e2.lo > ws ? e2.lo : ws, pb = e2.hi < we ? e2.hi : we;
                const double thr = e2.thr, xl = e2.xl, inv_dw = e2.inv_dw, eb = e2.b, ec = e2.c, Kc = e2.Kc, Kf = e2.Kf;
                if (pb - pa <= 64) {
                    for (int64_t c0 = pa; c0 < pb; c0 += 32) {
                        const int64_t pix = c0 + lane;
                        if (pix < pb) {
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

    // ---- far field: one polynomial per hierarchy level (Horner in t = (nu - nu_c) / h of that level's tile) -------
    if (use_far && warp_has_pixels) {
        constexpr int K1 = SD_FAR_K + 1;
#pragma unroll
        for (int lev = 0; lev < SD_FAR_LEVELS; lev++) {
            const int tk = tile >> (SD_FAR_SHIFT * lev);
            const double nu_c = a.fg.geom[lev][2 * tk], h_k = a.fg.geom[lev][2 * tk + 1];
            const double inv_h = (h_k > 0.0) ? 1.0 / h_k : 0.0;  // a one-pixel tile has h = 0: nothing is ever far from it, C = 0
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
    else k_lines<P, NW, false, 2, MINB><<<grid, 32 * NW, 0, c->stream>>>(a);
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
        // Small level-0 tiles keep the directly evaluated near field small (the hierarchy absorbs the rest), wide
        // per-warp spans keep the per-pair staging amortised: 2 warps x 256 pixels (1 warp on small grids).
        P = 8;
        NW = (((W + 511) / 512) * c->D >= 2LL * c->sm_count) ? 2 : 1;
    } else if (((W + 2047) / 2048) * c->D >= 8LL * c->sm_count) P = 8;
    else if (((W + 1023) / 1024) * c->D >= 4LL * c->sm_count) P = 4;
    else if (((W + 511) / 512) * c->D >= 4LL * c->sm_count) P = 2;
    else P = 1;
    if (force_p == 1 || force_p == 2 || force_p == 4 || force_p == 8) P = force_p;
    if (force_nw == 1 || force_nw == 2 || force_nw == 8) NW = force_nw;
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
    a.out = c->alpha_line[slot].as<double>();
    a.stats = c->stats.as<unsigned long long>();
    if (c->farfield) {
        for (int k = SD_FAR_LEVELS - 1; k >= 0; k--) {
            const int tk = a.fg.tile[k];
            a.far_tile0[k] = (int)(c->p0 / tk);
            a.far_ntl[k] = (int)((c->p1 + tk - 1) / tk) - a.far_tile0[k];
            SD_TRY(sd_ensure(c, c->far_coef[k], sizeof(double) * c->D * a.far_ntl[k] * (SD_FAR_K + 1)));
            a.far_coef[k] = c->far_coef[k].as<double>();
        }
        // CTAs per (group of eight sibling tiles, depth): the candidate lists are cut into this many fixed slices so that
        // even a narrow shard fills the chip; k_far_reduce adds the partial sums in slice order.
        if (!c->far_attr_set) {  // per device: > 48 KB of dynamic shared memory needs the opt-in; series-length table
            SD_CUDA(c, cudaFuncSetAttribute(k_far_coeffs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAR_SMEM));
            unsigned char tab[256];
            far_terms_table(tab);
            SD_CUDA(c, cudaMemcpyToSymbolAsync(FAR_TERMS, tab, sizeof tab, 0, cudaMemcpyHostToDevice, c->stream));
            SD_CUDA(c, cudaStreamSynchronize(c->stream));  // `tab` lives on this stack frame
            c->far_attr_set = true;
        }
        size_t part_bytes = 0;
        for (int k = 0; k < SD_FAR_LEVELS; k++) {
            const size_t b = sizeof(double) * c->D * a.far_ntl[k] * far_nsplit(k) * (SD_FAR_K + 1);
            part_bytes = b > part_bytes ? b : part_bytes;
        }
        SD_TRY(sd_ensure(c, c->far_part, part_bytes));
        sd_phase_begin(c, SD_PH_FAR);
        for (int k = SD_FAR_LEVELS - 1; k >= 0; k--) {
            const int nsplit = far_nsplit(k);
            // one CTA per group of eight sibling tiles that has a member in the launched range, times the slices
            const int n_cta = (((a.far_tile0[k] + a.far_ntl[k] - 1) >> SD_FAR_SHIFT) - (a.far_tile0[k] >> SD_FAR_SHIFT) + 1) * nsplit;
            k_far_coeffs<<<dim3((unsigned)n_cta, (unsigned)c->D), THREADS, FAR_SMEM, c->stream>>>(
                a, k, c->line_stats ? 1 : 0, nsplit, c->far_part.as<double>());
            SD_TRY(sd_launch_check(c, "k_far_coeffs"));
            if (nsplit > 1) {
                const int n = c->D * a.far_ntl[k] * (SD_FAR_K + 1);
                k_far_reduce<<<(n + 255) / 256, 256, 0, c->stream>>>(n, nsplit, c->far_part.as<double>(), a.far_coef[k]);
                SD_TRY(sd_launch_check(c, "k_far_reduce"));
            }
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
