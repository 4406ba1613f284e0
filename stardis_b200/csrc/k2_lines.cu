// k2_lines.cu -- K2: windowed Voigt accumulation  alpha_line[d, i] = sum_l phi(nu_i - nu_l; dw[l,d], gamma[l,d]) * alpha[l,d]
// over the pixels i in the (line, depth) window [lo, hi) only.
//
// Reference: stardis/radiation_field/opacities/opacities_solvers/base.py:487-627 (calc_alan_entries, thread-
// parallel over lines with one private (D,N) slab per numba thread) and voigt.py:17-155.
//
// B200 design (gather, no atomics, deterministic):
//   * one CTA owns a tile of 256*P consecutive pixels of ONE depth point and keeps the P accumulators of
//     every thread in registers for the whole kernel; the result is written exactly once, coalesced;
//   * the (line, depth) pairs that can touch the tile are found per half-width class (class 0: contiguous
//     range of the nu-sorted line list; class k >= 1: contiguous range of the per-depth class list built by
//     k1_broadening.cu; 32-ary warp binary searches on the monotone window centres);
//   * every WARP streams the candidates in batches of 32 on its own (no CTA barrier in the main loop): test against
//     the warp's 32*P-pixel span, expand the passing (line, depth) records into 96-byte shared-memory entries
//     (constants hoisted once per warp), entries whose window covers the span and whose span lies entirely in
//     Humlicek region I packed first ("far" list), the others from the back ("mixed" list), then consume them with
//     broadcast LDS; summation order per pixel is fixed (class, batch, far entries in line order, mixed reversed);
//   * the hot loop is the far-wing (region I) form  Kf (q + c1) / (q (q + b) + c),  q = x^2:
//     8 FP64 instructions + 1 MUFU.RCP64H per evaluation (x, q, 2 for the denominator, numerator, 2 for the Newton
//     step on the reciprocal seed, accumulate), no branches, no divisions;
//   * pixels that are not certainly in region I take the exact path: x = dnu / dw (IEEE division) and the
//     reference's own region tests, so the Humlicek classification is identical to the reference's.
//
// Roofline: FP64 FMA pipe (no dense contraction -> no tensor cores).  Memory traffic is negligible:
// 64 B per candidate record per tile, 8 B per output cell.
#include <stdlib.h>

#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
static_assert(WARPS == SD_NCLS, "one warp per half-width class in the range search");

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

__device__ __noinline__ double exact_contribution(double nu_i, double nu_l, double dw, double y, double K) {
    double x = (nu_i - nu_l) / dw;  // voigt.py:148, IEEE division
    return sdm::humlicek_re(x, y) * K;
}

template <int P, bool STATS, int RCP, bool U2, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_lines(int64_t L, int D, int64_t N, int64_t p0, int64_t p1,
                                                   const double *__restrict__ nus, const int *__restrict__ line_idx,
                                                   const LineRec *__restrict__ rec, const int *__restrict__ win_lo,
                                                   const int *__restrict__ win_hi, const uint8_t *__restrict__ win_cls,
                                                   const int *__restrict__ cls_list, const int *__restrict__ cls_off,
                                                   double *__restrict__ out, unsigned long long *__restrict__ stats) {
    constexpr int TILE = THREADS * P;
    constexpr int SPAN = 32 * P;
    __shared__ WEntry s_ent[WARPS][32];  // every warp streams its own batches: no CTA barrier in the main loop
    __shared__ int s_ja[SD_NCLS], s_jb[SD_NCLS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int d = blockIdx.y;
    const int64_t t0 = p0 + (int64_t)blockIdx.x * TILE;
    const int64_t t1 = (t0 + TILE < p1) ? t0 + TILE : p1;
    const int64_t ws = t0 + (int64_t)warp * SPAN;                 // first pixel of this warp's span
    const int64_t we = (ws + SPAN < t1) ? ws + SPAN : t1;         // one past its last valid pixel
    const size_t drow = (size_t)d * L;
    const int *list_d = cls_list + drow;

    // pixel frequencies and accumulators live in registers for the whole kernel
    double nu_i[P], acc[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        nu_i[p] = nus[pix < N ? pix : N - 1];
        acc[p] = 0.0;
    }
    // frequencies at the two ends of this warp's span (x is monotone in the pixel index)
    const double nu_first = nus[ws < N ? ws : N - 1];
    const double nu_last = nus[(we - 1 >= ws && we - 1 < N) ? we - 1 : (ws < N ? ws : N - 1)];

    // candidate ranges of all classes, one warp per class
    {
        const int cls = warp;  // WARPS == SD_NCLS
        int ja, jb;
        if (cls == 0) {
            auto key = [&](int j) { return line_idx[j]; };
            long long Xa = t1 + SD_CLS0_HW, Xb = t0 - SD_CLS0_HW + 1;  // idx < t1+H ; idx <= t0-H
            ja = warp_first_below(key, 0, (int)L, (int)(Xa > 2147483647LL ? 2147483647LL : Xa));
            jb = warp_first_below(key, ja, (int)L, (int)(Xb < -2147483647LL ? -2147483647LL : Xb));
        } else {
            int a = cls_off[d * (SD_NCLS + 1) + cls], b = cls_off[d * (SD_NCLS + 1) + cls + 1];
            if (cls == SD_NCLS - 1) {
                ja = a;
                jb = b;
            } else {
                long long H = (long long)SD_CLS0_HW << (2 * cls);
                auto key = [&](int j) { return line_idx[list_d[j]]; };
                long long Xa = t1 + H, Xb = t0 - H + 1;
                ja = warp_first_below(key, a, b, (int)(Xa > 2147483647LL ? 2147483647LL : Xa));
                jb = warp_first_below(key, ja, b, (int)(Xb < -2147483647LL ? -2147483647LL : Xb));
            }
        }
        if (lane == 0) {
            s_ja[cls] = ja;
            s_jb[cls] = jb;
        }
    }
    __syncthreads();

    unsigned long long h0 = 0, h1 = 0, h2 = 0, h3 = 0;
    int nvalid = 0;  // this lane's pixels inside the tile (statistics only)
#pragma unroll
    for (int p = 0; p < P; p++) nvalid += (ws + p * 32 + lane) < t1;

    WEntry *const my = s_ent[warp];
    const bool warp_has_pixels = ws < t1;

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

    for (int cls = 0; cls < SD_NCLS && warp_has_pixels; cls++) {
        const int ja = s_ja[cls], jb = s_jb[cls];
        for (int base = ja; base < jb; base += 32) {
            // ---- test the 32 candidates of this batch against THIS WARP's span ------------------------
            const int j = base + lane;
            bool pass = false;
            size_t o = 0;
            int lo = 0, hi = 0;
            if (j < jb) {
                int l = (cls == 0) ? j : list_d[j];
                o = drow + l;
                lo = win_lo[o];
                hi = win_hi[o];
                pass = (lo < we) && (hi > ws) && (hi > lo) && (cls != 0 || win_cls[o] == 0);
            }
            if (!__any_sync(0xffffffffu, pass)) continue;
            // ---- stage: hoist the per-(line, depth) constants; far entries first, mixed ones from the back -----
            WEntry e;
            bool ff = false;
            if (pass) {
                const LineRec r = rec[o];
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
            // ---- consume: far-wing entries, two per iteration (independent chains, loads hoisted) ----------
            int k = 0;
            for (; U2 && k + 1 < n_far; k += 2) {
                const WEntry &a = my[k], &b = my[k + 1];
                const double a0 = a.xl, a1 = a.inv_dw, a2 = a.b, a3 = a.c, a4 = a.Kc, a5 = a.Kf;
                const double b0 = b.xl, b1 = b.inv_dw, b2 = b.b, b3 = b.c, b4 = b.Kc, b5 = b.Kf;
                far_eval(a0, a1, a2, a3, a4, a5);
                far_eval(b0, b1, b2, b3, b4, b5);
            }
            for (; k < n_far; k++) {
                const WEntry &a = my[k];
                far_eval(a.xl, a.inv_dw, a.b, a.c, a.Kc, a.Kf);
            }
            if (STATS) h0 += (unsigned long long)n_far * nvalid;
            // ---- mixed entries: window edge inside the span and/or pixels near the line core --------------
            for (int m = 0; m < n_mix; m++) {
                const WEntry &e2 = my[31 - m];
                const int lo2 = e2.lo, hi2 = e2.hi;
                const double thr = e2.thr, xl = e2.xl, inv_dw = e2.inv_dw, eb = e2.b, ec = e2.c, Kc = e2.Kc, Kf = e2.Kf;
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
                    if (inwin && !fast) acc[p] += exact_contribution(nu_i[p], e2.nu, e2.dw, e2.y, e2.K);
                    if (STATS && inwin) {
                        int r = sdm::humlicek_region((nu_i[p] - e2.nu) / e2.dw, e2.y);
                        h0 += (r == 0); h1 += (r == 1); h2 += (r == 2); h3 += (r == 3);
                    }
                }
            }
            __syncwarp();
        }
    }

#pragma unroll
    for (int p = 0; p < P; p++) {
        int64_t pix = ws + p * 32 + lane;
        if (pix < t1) out[(size_t)d * (p1 - p0) + (pix - p0)] = acc[p];
    }
    if (STATS) {
        for (int o2 = 16; o2; o2 >>= 1) {
            h0 += __shfl_xor_sync(0xffffffffu, h0, o2);
            h1 += __shfl_xor_sync(0xffffffffu, h1, o2);
            h2 += __shfl_xor_sync(0xffffffffu, h2, o2);
            h3 += __shfl_xor_sync(0xffffffffu, h3, o2);
        }
        if (lane == 0) {
            if (h0) atomicAdd(&stats[0], h0);
            if (h1) atomicAdd(&stats[1], h1);
            if (h2) atomicAdd(&stats[2], h2);
            if (h3) atomicAdd(&stats[3], h3);
        }
    }
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int P>
int launch(sd_ctx *c, int slot, bool stats, int rcp) {
    int64_t W = c->W();
    dim3 grid((unsigned)((W + THREADS * P - 1) / (THREADS * P)), (unsigned)c->D);
    auto args = [&](auto kern) {
        kern<<<grid, THREADS, 0, c->stream>>>(c->L, c->D, c->N, c->p0, c->p1, c->nus.as<double>(), c->line_idx.as<int>(),
                                              c->rec.as<LineRec>(), c->win_lo.as<int>(), c->win_hi.as<int>(),
                                              c->win_cls.as<uint8_t>(), c->cls_list.as<int>(), c->cls_off.as<int>(),
                                              c->alpha_line[slot].as<double>(), c->stats.as<unsigned long long>());
    };
    // the counting instantiation uses the production arithmetic (Newton reciprocal) so that both are bitwise equal
    static const int u2 = env_int("SD_K2_U2", 0);
    static const int minb = env_int("SD_K2_MINB", 2);
    if (stats) args(k_lines<P, true, 2, false, 2>);
    else if (rcp == 3) args(k_lines<P, false, 3, false, 2>);
    else if (u2) args(k_lines<P, false, 2, true, 2>);
    else if (minb == 1) args(k_lines<P, false, 2, false, 1>);
    else if (minb == 3) args(k_lines<P, false, 2, false, 3>);
    else if (minb == 4) args(k_lines<P, false, 2, false, 4>);
    else args(k_lines<P, false, 2, false, 2>);
    return sd_launch_check(c, "k_lines");
}

}  // namespace

int sd_k2_lines(sd_ctx *c, int slot) {
    int64_t W = c->W();
    SD_TRY(sd_ensure(c, c->alpha_line[slot], sizeof(double) * c->D * W));
    if (c->L == 0) {
        SD_CUDA(c, cudaMemsetAsync(c->alpha_line[slot].p, 0, sizeof(double) * c->D * W, c->stream));
        return SD_OK;
    }
    if (c->line_stats) SD_CUDA(c, cudaMemsetAsync(c->stats.p, 0, 4 * sizeof(unsigned long long), c->stream));
    // pixels per thread: enough CTAs to fill the chip several times over, otherwise as much register reuse
    // of the staged entries as possible.  SD_K2_P / SD_K2_RCP override the choice (tuning experiments).
    static const int force_p = env_int("SD_K2_P", 0);
    static const int rcp = env_int("SD_K2_RCP", 2);
    int P = 1;
    if (((W + 2047) / 2048) * c->D >= 8LL * c->sm_count) P = 8;
    else if (((W + 1023) / 1024) * c->D >= 4LL * c->sm_count) P = 4;
    else if (((W + 511) / 512) * c->D >= 4LL * c->sm_count) P = 2;
    if (force_p) P = force_p;
    switch (P) {
        case 8: return launch<8>(c, slot, c->line_stats, rcp);
        case 4: return launch<4>(c, slot, c->line_stats, rcp);
        case 2: return launch<2>(c, slot, c->line_stats, rcp);
        default: return launch<1>(c, slot, c->line_stats, rcp);
    }
}
