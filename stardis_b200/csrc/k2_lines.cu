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
//   * candidates are tested against the tile, compacted in line order, expanded into 112-byte shared-memory
//     entries (per-(line,depth) constants hoisted once per CTA, incl. per-warp "fully inside the window and
//     entirely in Humlicek region I" flags) and then consumed by all 8 warps with broadcast LDS;
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
constexpr int STAGE = 256;  // entries staged per round (one candidate per thread)
static_assert(WARPS == SD_NCLS, "one warp per half-width class in the range search");

struct __align__(16) SEntry {
    // fast path (64 B)
    double xl;      // nu_l / dw
    double inv_dw;  // 1 / dw
    double thr;     // q > thr  =>  region I for certain
    double b;       // 2 y^2 - 1
    double c;       // (y^2 + 1/2)^2
    double Kc;      // Kf (y^2 + 1/2)
    double Kf;      // alpha y / (pi dw)
    int lo, hi;     // window
    // exact path (32 B)
    double nu, dw, y, K;
    // per-warp flags: bit w = warp w's span overlaps the window / lies fully inside it and fully in region I
    unsigned m_overlap, m_fullfar;
    unsigned pad0, pad1;
};
static_assert(sizeof(SEntry) == 112, "SEntry layout");

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

template <int P, bool STATS, int RCP>
__global__ void __launch_bounds__(THREADS) k_lines(int64_t L, int D, int64_t N, int64_t p0, int64_t p1,
                                                   const double *__restrict__ nus, const int *__restrict__ line_idx,
                                                   const LineRec *__restrict__ rec, const int *__restrict__ win_lo,
                                                   const int *__restrict__ win_hi, const uint8_t *__restrict__ win_cls,
                                                   const int *__restrict__ cls_list, const int *__restrict__ cls_off,
                                                   double *__restrict__ out, unsigned long long *__restrict__ stats) {
    constexpr int TILE = THREADS * P;
    constexpr int SPAN = 32 * P;
    __shared__ SEntry s_ent[STAGE];
    __shared__ double s_edge[WARPS][2];
    __shared__ int s_ja[SD_NCLS], s_jb[SD_NCLS];
    __shared__ int s_wcnt[WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    if (lane == 0) {
        int64_t a = ws < N ? ws : N - 1, b = (we - 1 >= ws) ? we - 1 : a;
        s_edge[warp][0] = nus[a];
        s_edge[warp][1] = nus[b < N ? b : N - 1];
    }

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

    for (int cls = 0; cls < SD_NCLS; cls++) {
        const int ja = s_ja[cls], jb = s_jb[cls];
        for (int base = ja; base < jb; base += STAGE) {
            // ---- test + ordered compaction ---------------------------------------------------------
            int j = base + tid;
            bool pass = false;
            size_t o = 0;
            int lo = 0, hi = 0;
            if (j < jb) {
                int l = (cls == 0) ? j : list_d[j];
                o = drow + l;
                lo = win_lo[o];
                hi = win_hi[o];
                pass = (lo < t1) && (hi > t0) && (hi > lo) && (cls != 0 || win_cls[o] == 0);
            }
            unsigned bal = __ballot_sync(0xffffffffu, pass);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            int pos = __popc(bal & ((1u << lane) - 1u));
            int total = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                int cw = s_wcnt[w];
                if (w < warp) pos += cw;
                total += cw;
            }
            // ---- stage: hoist the per-(line, depth) constants once per CTA ------------------------
            if (pass) {
                const LineRec r = rec[o];
                SEntry e;
                double yy = r.y * r.y;
                e.xl = r.nu * r.inv_dw;
                e.inv_dw = r.inv_dw;
                e.thr = r.thr;
                const double c1 = yy + 0.5;
                e.b = 2.0 * yy - 1.0;
                e.c = c1 * c1;
                e.Kf = r.K * r.y * sdm::INV_SQRT_PI;
                e.Kc = e.Kf * c1;
                e.lo = lo;
                e.hi = hi;
                e.nu = r.nu;
                e.dw = r.dw;
                e.y = r.y;
                e.K = r.K;
                unsigned mo = 0, mf = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++) {
                    int64_t a = t0 + (int64_t)w * SPAN;
                    int64_t b = (a + SPAN < t1) ? a + SPAN : t1;
                    if (a < b && lo < b && hi > a) {
                        mo |= 1u << w;
                        if (lo <= a && hi >= b) {
                            double xa = fma(s_edge[w][0], e.inv_dw, -e.xl);
                            double xb = fma(s_edge[w][1], e.inv_dw, -e.xl);
                            if (xa * xb > 0.0 && fmin(xa * xa, xb * xb) > e.thr) mf |= 1u << w;
                        }
                    }
                }
                e.m_overlap = mo;
                e.m_fullfar = mf;
                e.pad0 = e.pad1 = 0;
                s_ent[pos] = e;
            }
            __syncthreads();
            // ---- consume: every warp walks the staged entries for its own pixel span ---------------
            for (int k = 0; k < total; k++) {
                const SEntry &e = s_ent[k];
                const unsigned mo = e.m_overlap, mf = e.m_fullfar;
                if (!((mo >> warp) & 1u)) continue;
                const double xl = e.xl, inv_dw = e.inv_dw, eb = e.b, ec = e.c, Kc = e.Kc, Kf = e.Kf;
                if ((mf >> warp) & 1u) {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        double x = fma(nu_i[p], inv_dw, -xl);
                        double q = x * x;
                        double den = fma(q, q + eb, ec);
                        double num = fma(Kf, q, Kc);
                        acc[p] = fma(num, RCP == 2 ? sdm::rcp_fast2(den) : sdm::rcp_fast(den), acc[p]);
                    }
                    if (STATS) h0 += nvalid;
                } else {
                    const int lo2 = e.lo, hi2 = e.hi;
                    const double thr = e.thr;
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
                        if (inwin && !fast) acc[p] += exact_contribution(nu_i[p], e.nu, e.dw, e.y, e.K);
                        if (STATS && inwin) {
                            int r = sdm::humlicek_region((nu_i[p] - e.nu) / e.dw, e.y);
                            h0 += (r == 0); h1 += (r == 1); h2 += (r == 2); h3 += (r == 3);
                        }
                    }
                }
            }
            __syncthreads();
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
    if (stats) args(k_lines<P, true, 2>);
    else if (rcp == 3) args(k_lines<P, false, 3>);
    else args(k_lines<P, false, 2>);
    return sd_launch_check(c, "k_lines");
}

int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
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
