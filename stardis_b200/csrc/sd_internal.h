// sd_internal.h -- context object and helpers shared by the translation units of libstardis_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/stardis_b200.h"

// ---- K2 line-record layout (device global memory, depth-major: rec[d * L + l]) ----------------------
struct __align__(16) LineRec {
    // first 32 bytes: everything the far-field expansion needs (staged with two 16-byte async copies)
    double nu;      // line frequency
    double dw;      // doppler width (exact division in the near-core path)
    double y;       // (gamma / (sqrt(pi) pi)) / dw            voigt.py:148
    double K;       // alpha_line / (sqrt(pi) dw)              voigt.py:149, base.py:627
    double inv_dw;  // 1 / doppler width
    double thr;     // x^2 above which the pixel is certainly Humlicek region I (or -1: always)
    double pad0, pad1;
};
static_assert(sizeof(LineRec) == 64, "LineRec must be 64 bytes");

// ---- per-(depth, line) window record: everything the candidate tests of k_lines / k_far_coeffs / k_s2m need, in ONE
// 16-byte load per candidate
struct __align__(16) PairWin {
    int lo, hi;         // window [lo, hi) in global pixels (base.py:563-592 semantics, see sdm::line_window)
    int cpix;           // pixel of the line centre (line_idx clamped to the grid): tile of level k = cpix >> tile_shift[k]
    unsigned char cls;  // half-width class (0..6) or SD_FC0 + lmin for far-capable pairs
    unsigned char lmin; // lowest hierarchy level at which the pair may be expanded (SD_FAR_LEVELS: never)
    unsigned char sat;  // bit k: the window covers the whole interaction neighbourhood of the pair's level-k tile
    unsigned char pad;
};
static_assert(sizeof(PairWin) == 16, "PairWin must be 16 bytes");

constexpr int SD_CLS0_HW = 64;      // class 0: hw <= 64; class k: hw <= 64 * 4^k; class 6: everything wider
constexpr int SD_MAX_SOURCES = 4 + SD_MAX_TABLES;
// Far field of the region-I wings on a hierarchy of pixel tiles (64 / 512 / 4096 / 32768 pixels, branching 8).
// A (line, depth) pair whose centre lies in tile s of level k is FAR from tile t of that level when |s - t| >= 2, its
// window covers t, t has at least two pixels and k >= lmin(pair) (lmin: the lowest level at which every pixel two tiles
// away is certainly in Humlicek region I and the pair's two poles sit close enough to their own tile for the
// expansions below to converge with ratio <= SD_FAR_RHO).  A far (pair, pixel) product is served at the HIGHEST level
// at which it is far, by a polynomial of degree SD_FAR_K in (nu - nu_c) / h per tile and depth:
//   * pairs whose window covers the whole interaction neighbourhood of their tile ("saturated": the 24 children of
//     the parent tile and its two neighbours; the whole grid at the top level) go through multipole moments of their
//     own tile (k_s2m, ONE expansion per pair and level) and tile-to-tile translations (k_m2l), as in a 1-D fast
//     multipole method -- the translation matrices are real because the tile centres are;
//   * the others (a window edge inside the neighbourhood) are expanded directly about the target tiles
//     (k_far_coeffs), as every far pair was in round 1.
constexpr int SD_FAR_K = 31;             // polynomial degree (32 coefficients); multipole moments k = 1..32
constexpr double SD_FAR_RHO = 0.40;      // largest admitted convergence ratio: 33 * 0.4^32 = 6e-12
constexpr float SD_FAR_LOG2_RHO_INV = 1.3219281f;  // log2(1 / SD_FAR_RHO)
constexpr double SD_FAR_OVERHANG = 0.2;  // the poles may leave their tile by at most this fraction of its half-width
constexpr int SD_FAR_LEVELS = 4;         // tile hierarchy: level k tiles hold 64 * 8^k pixels
constexpr int SD_FAR_SHIFT = 3;          // log2 of the branching factor
constexpr int SD_FAR_TILE0_SHIFT = 6;    // level-0 tiles: 64 pixels = two register slots of a k_lines warp
constexpr int SD_FC0 = 7;                // classes SD_FC0 + m: far-capable pairs with lmin = m (sorted by centre)
constexpr int SD_NCLS = SD_FC0 + SD_FAR_LEVELS;  // half-width classes 0..6 + far-capable classes

// geometry of the far-field tile hierarchy and the per-pair tables that drive it, passed by value to the kernels
struct FarGeom {
    int tile[SD_FAR_LEVELS];            // pixels per tile (powers of two)
    int tile_shift[SD_FAR_LEVELS];      // log2 of them
    int n_tiles[SD_FAR_LEVELS];         // global number of tiles
    const double *geom[SD_FAR_LEVELS];  // {centre frequency, half-width, moment scale} per tile (3 doubles)
    int enabled;                        // far field on (far-capable classes and edge lists exist)
    // [SD_FAR_LEVELS + 1] written by k_level_check: n_active (levels 0 .. n_active-1 are usable on this grid; level
    // n_active-1 is the top level, whose interaction neighbourhood is the whole grid), then per level the admitted
    // pole overhang as a fraction of the tile half-width (float bits)
    const int *lev_info;
    // Window edges of the far-capable pairs that lie inside the grid AND inside this context's extended pixel range
    // (the top-level tiles its range touches), as ONE sorted array of 64-bit keys
    //   kind (0 = window start, 1 = window end) | depth | lmin | edge pixel | line index      (see sd_edge_key)
    // -- a total order, so the result does not depend on the order in which k_build_records appended them.
    // edge_off[(kind * D + d) * SD_FAR_LEVELS + m] = first entry of (kind, d, m); one more entry closes the array.
    const unsigned long long *edge_keys;
    const int *edge_off;
    // Range tables per level-0 tile boundary t = 0..n_tiles[0] (k_range_tables): the far-capable pairs of (depth d,
    // lmin m) centred in the level-0 tiles [ta, tb) are the entries [fc_tab[row + tb], fc_tab[row + ta]) of the depth's
    // class list (centres descend along the list), row = (d * SD_FAR_LEVELS + m) * (n_tiles[0] + 1); their window edges
    // at pixels [64 ta, 64 tb) are the keys [edge_tab[erow + ta], edge_tab[erow + tb)), erow = ((kind * D + d) *
    // SD_FAR_LEVELS + m) * (n_tiles[0] + 1).  Two loads instead of two binary searches per candidate range.
    const int *fc_tab;
    const int *edge_tab;
    int l_bits, pix_bits, depth_bits;
    unsigned long long *edge_out;        // unsorted append buffer (k_build_records)
    unsigned long long *edge_count;      // [1] number of appended keys
    long long ext0, ext1;                // extended pixel range [ext0, ext1): pairs whose window misses it get no LineRec
};
constexpr int SD_FAR_LMIN_BITS = 2;
static_assert((1 << SD_FAR_LMIN_BITS) >= SD_FAR_LEVELS, "lmin field of the edge key");

__host__ __device__ __forceinline__ unsigned long long sd_edge_key(const FarGeom &fg, int kind, int d, int lmin, long long pixel,
                                                                   int l) {
    return (((((unsigned long long)kind << fg.depth_bits | (unsigned long long)d) << SD_FAR_LMIN_BITS | (unsigned long long)lmin)
             << fg.pix_bits | (unsigned long long)pixel) << fg.l_bits) | (unsigned long long)(unsigned)l;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

// device-side phase timers: CUDA events recorded on the context's stream around each group of kernels, read back by
// sd_phase_times (bench.py's per-kernel roofline figures are measured live with these, not under a profiler)
constexpr int SD_N_PHASES = 8;
enum { SD_PH_K1 = 0, SD_PH_PREP = 1, SD_PH_SORT = 2, SD_PH_FAR = 3, SD_PH_LINES = 4, SD_PH_K3 = 5, SD_PH_K4 = 6, SD_PH_STRENGTH = 7 };
constexpr int SD_N_STATS = 16;  // uint64 counters, see sd_line_stats_ex

struct sd_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;      // stream in use
    cudaStream_t own_stream = nullptr;  // created by sd_create
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int sm_count = 148;
    int64_t launches = 0;  // kernels launched so far
    cudaEvent_t ph_ev[SD_N_PHASES][2] = {};
    bool ph_rec[SD_N_PHASES] = {};

    // atmosphere
    int D = 0;
    double vmic = 0.0;
    DevBuf T, ne, nH;

    // grid
    int64_t N = 0, p0 = 0, p1 = 0;
    DevBuf nus;   // [N]
    DevBuf d_nu;  // [1]

    // lines
    int64_t L = 0;
    bool has_atomic_cols = false, has_vald_cols = false;
    DevBuf l_nu, l_Z, l_ion, l_eion, l_eup, l_elo, l_A, l_mass, l_stark, l_waals, l_alpha;
    DevBuf gammas, dws;
    DevBuf vald_stage;            // staged inputs of sd_calc_alpha_line_vald
    bool have_alpha_line = false; // l_alpha holds line strengths (uploaded or computed on the device)
    int gamma_cols = 0;
    bool have_broadening = false;

    // K2 preparation
    DevBuf line_idx;   // int32 [L]
    DevBuf rec;        // LineRec [D*L]
    DevBuf win;        // PairWin [D*L]
    DevBuf win_cls;    // uint8 [D*L]
    DevBuf cls_list;   // int32 [D*L]   per depth: lines of class 1.. concatenated by class (stable in l)
    DevBuf cls_off;    // int32 [D*(NCLS+1)] offsets into cls_list row d (class 0 is not listed)
    DevBuf chunk_cnt;  // int32 [D * nchunks * NCLS]
    DevBuf stats;      // uint64 [8]
    DevBuf lev_info;                   // int [SD_FAR_LEVELS + 1], see FarGeom::lev_info
    int far_active = 0;                // host copy of lev_info[0] (read at the edge-sort synchronisation)
    DevBuf edge_keys, edge_unsorted;   // 64-bit window-edge keys (sorted / as appended), see FarGeom
    DevBuf edge_off, edge_count, edge_sort_tmp;
    DevBuf fc_tab, edge_tab;           // range tables per level-0 tile boundary, see FarGeom
    DevBuf line_pre, depth_pre;        // K1: per-line / per-depth factors of the broadening formulae (pow() hoisted)
    unsigned long long *h_edge_count = nullptr;  // pinned host copy of the edge counter
    DevBuf tile_geom[SD_FAR_LEVELS];   // double [3 * n_tiles]: centre frequency, half-width, moment scale of every global tile
    DevBuf far_coef[SD_FAR_LEVELS];    // double [D * n_tiles_shard * (SD_FAR_K + 1)]
    DevBuf far_mom[SD_FAR_LEVELS];     // double [D * n_tiles * (SD_FAR_K + 1)]: multipole moments of the saturated pairs
    DevBuf far_bkt[SD_FAR_LEVELS][SD_FAR_LEVELS];  // [level][highest saturated level]: moment buckets, see k_s2m / k_m2m
    DevBuf far_part;                   // partial coefficient sets (slices of the pair list per tile)
    FarGeom far_geom{};
    int k2_P = 4;      // pixels per thread chosen for the current grid
    int k2_NW = 8;     // warps per CTA of the line kernel (CTA tile = 32 * k2_NW * k2_P pixels)
    bool farfield = true;
    bool far_attr_set = false;  // dynamic shared memory opt-in of k_far_coeffs done on this device
    DevBuf alpha_line[2];
    bool have_alpha[2] = {false, false};
    bool records_ready = false;
    bool line_stats = false;  // run the statistics-collecting instantiation of k_lines

    // K3 / K4
    DevBuf total;
    bool have_total = false;
    DevBuf src[SD_MAX_SOURCES];
    bool have_src[SD_MAX_SOURCES] = {};
    DevBuf cont_small;  // packed small arrays of the continuum descriptor
    DevBuf F, I_nus, ray_small;
    bool have_F = false;
    int n_theta_tracked = 0;

    int64_t W() const { return p1 - p0; }
};

int sd_fail(sd_ctx *c, int code, const char *fmt, ...);

inline void sd_phase_begin(sd_ctx *c, int ph) { cudaEventRecord(c->ph_ev[ph][0], c->stream); }
inline void sd_phase_end(sd_ctx *c, int ph) {
    cudaEventRecord(c->ph_ev[ph][1], c->stream);
    c->ph_rec[ph] = true;
}

#define SD_CUDA(c, call)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return sd_fail((c), SD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                       \
    } while (0)

#define SD_CHECK(c, cond, code, ...) \
    do {                             \
        if (!(cond)) return sd_fail((c), (code), __VA_ARGS__); \
    } while (0)

#define SD_TRY(expr)          \
    do {                      \
        int _r = (expr);      \
        if (_r != SD_OK) return _r; \
    } while (0)

// grow-only device allocation / transfers (host or device source, UVA)
int sd_ensure(sd_ctx *c, DevBuf &b, size_t bytes);
int sd_upload(sd_ctx *c, DevBuf &b, const void *src, size_t bytes);
int sd_launch_check(sd_ctx *c, const char *what);

// kernel launchers implemented in the other translation units
int sd_k1_broadening(sd_ctx *c, uint32_t flags);
int sd_k2_prepare(sd_ctx *c);
int sd_k2_lines(sd_ctx *c, int slot);
int sd_k2_choose_P(sd_ctx *c);
int sd_sort_edges(sd_ctx *c);  // k2_sort.cu (CUB radix sort of the appended window-edge keys + segment offsets)
int sd_k3_continuum(sd_ctx *c, const sd_continuum *desc, uint32_t store_mask);
int sd_k4_raytrace(sd_ctx *c, int n_theta, const double *ray_ds, const double *weights, int inward, double scale,
                   int track);
