// k3_continuum.cu -- K3: all continuum opacity terms in ONE depth x nu pass, plus the total.
//
// Reference (stardis/radiation_field/opacities/opacities_solvers/base.py): calc_alpha_file :40-70 with
// sigma_file (util.py:14-108), calc_alpha_bf :178-271, calc_alpha_ff :274-317, calc_alpha_rayleigh :74-135,
// calc_alpha_electron :139-174, and Opacities.calc_total_alphas (opacities/base.py:24-28).  The reference
// materialises one (D,N) array per term (bf with a Python loop over every nu); here one thread owns one
// frequency, hoists everything that only depends on nu (nu^-3, Rayleigh powers, bound-free edge index,
// table cell + fraction) and walks the depth points, adding the line opacity slots and writing the total
// (and, on request, each term) coalesced.
//
// Roofline: HBM.  Algorithmic bytes per (depth, nu) cell: 8 B read per computed alpha_line slot + 8 B write
// of the total (+ 8 B per separately stored term).
#include <vector>

#include "sd_internal.h"
#include "sd_math.cuh"

namespace {

struct TableDev {
    int kind, nx, ny;
    const double *x, *y, *values, *depth_y, *depth_scale;
    const uint8_t *diag;
};

struct ContDev {
    int n_bf;
    const double *bf_nu_cut, *bf_prefix;  // (n_bf+1, D)
    const double *ff_coef, *c4, *c6, *c8, *electron;
    int n_tables;
    TableDev tab[SD_MAX_TABLES];
    double *src_out[SD_MAX_SOURCES];  // nullptr = do not store
    const double *line0, *line1;      // alpha_line slots or nullptr
};

// largest j with a[j] <= v in ascending a[0..n) (n >= 1, caller guarantees a[0] <= v)
__device__ __forceinline__ int cell_of(const double *__restrict__ a, int n, double v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (a[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_continuum(int D, int64_t N, int64_t p0, int64_t p1, const double *__restrict__ nus,
                                                   ContDev cd, double *__restrict__ total) {
    int64_t i = p0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p1) return;
    const int64_t W = p1 - p0, col = i - p0;
    const double nu = nus[i];
    const double inv_nu3 = pow(nu, -3.0);  // base.py:206, 297
    // Rayleigh: frequencies above 2.3e15 Hz count as 0 (base.py:98-99)
    const double nu_r = (nu > 2.3e15) ? 0.0 : nu;
    const double rr = nu_r / (2.0 * (sdm::C_CGS * sdm::RYD_CGS));
    const double r2 = rr * rr, r4 = r2 * r2, r6 = r4 * r2, r8 = r4 * r4;
    // bound-free: number of levels whose cutoff is <= nu (base.py:266 "nu >= cutoff_frequency")
    int kbf = 0;
    if (cd.n_bf > 0 && nu >= cd.bf_nu_cut[0]) kbf = cell_of(cd.bf_nu_cut, cd.n_bf, nu) + 1;
    // tables: wavelength cell and fraction
    const double lam = sdm::C_CGS / nu * 1e8;
    int tix[SD_MAX_TABLES];
    double tfx[SD_MAX_TABLES];
    bool tin[SD_MAX_TABLES];
#pragma unroll
    for (int t = 0; t < SD_MAX_TABLES; t++) {
        tix[t] = 0; tfx[t] = 0.0; tin[t] = false;
        if (t < cd.n_tables) {
            const TableDev &tb = cd.tab[t];
            if (tb.kind == 1) {  // np.interp: clamp to the end values (util.py:99-103)
                if (lam <= tb.x[0]) { tix[t] = 0; tfx[t] = 0.0; }
                else if (lam >= tb.x[tb.nx - 1]) { tix[t] = tb.nx - 1; tfx[t] = 0.0; }
                else { int j = cell_of(tb.x, tb.nx, lam); tix[t] = j; tfx[t] = lam - tb.x[j]; }
                tin[t] = true;
            } else {
                if (lam >= tb.x[0] && lam <= tb.x[tb.nx - 1]) {
                    int j = cell_of(tb.x, tb.nx, lam);
                    if (j == tb.nx - 1) j = tb.nx - 2;
                    tix[t] = j;
                    tfx[t] = (lam - tb.x[j]) / (tb.x[j + 1] - tb.x[j]);
                    tin[t] = true;
                }
            }
        }
    }

    for (int d = 0; d < D; d++) {
        double tot = 0.0;
        const size_t o = (size_t)d * W + col;
        // order of accumulation follows calc_alphas: file, bf, ff, rayleigh, electron, line, molecule
        for (int t = 0; t < cd.n_tables; t++) {
            const TableDev &tb = cd.tab[t];
            double v = 0.0;
            if (tb.kind == 1) {
                int j = tix[t];
                if (tfx[t] == 0.0) v = tb.values[j];
                else {
                    double slope = (tb.values[j + 1] - tb.values[j]) / (tb.x[j + 1] - tb.x[j]);
                    v = slope * tfx[t] + tb.values[j];
                }
            } else if (tin[t]) {
                double yq = tb.depth_y[d];
                if (yq >= tb.y[0] && yq <= tb.y[tb.ny - 1]) {
                    int jy = cell_of(tb.y, tb.ny, yq);
                    if (jy == tb.ny - 1) jy = tb.ny - 2;
                    double fy = (yq - tb.y[jy]) / (tb.y[jy + 1] - tb.y[jy]);
                    double fx = tfx[t];
                    int ix = tix[t];
                    double v00 = tb.values[ix * tb.ny + jy], v01 = tb.values[ix * tb.ny + jy + 1];
                    double v10 = tb.values[(ix + 1) * tb.ny + jy], v11 = tb.values[(ix + 1) * tb.ny + jy + 1];
                    if (tb.diag[ix * (tb.ny - 1) + jy] == 0) {  // diagonal (0,0)-(1,1)
                        v = (fx >= fy) ? v00 + fx * (v10 - v00) + fy * (v11 - v10)
                                       : v00 + fy * (v01 - v00) + fx * (v11 - v01);
                    } else {  // diagonal (1,0)-(0,1)
                        v = (fx + fy <= 1.0) ? v00 + fx * (v10 - v00) + fy * (v01 - v00)
                                             : v11 + (1.0 - fx) * (v01 - v11) + (1.0 - fy) * (v10 - v11);
                    }
                }
            }
            v *= tb.depth_scale[d];
            if (cd.src_out[SD_SRC_TABLE0 + t]) cd.src_out[SD_SRC_TABLE0 + t][o] = v;
            tot += v;
        }
        if (cd.bf_prefix) {
            double v = cd.bf_prefix[(size_t)kbf * D + d] * inv_nu3;
            if (cd.src_out[SD_SRC_BF]) cd.src_out[SD_SRC_BF][o] = v;
            tot += v;
        }
        if (cd.ff_coef) {
            double v = cd.ff_coef[d] * inv_nu3;
            if (cd.src_out[SD_SRC_FF]) cd.src_out[SD_SRC_FF][o] = v;
            tot += v;
        }
        if (cd.c4) {
            double v = (cd.c4[d] * r4 + cd.c6[d] * r6 + cd.c8[d] * r8) * 6.6524587321e-25;
            if (cd.src_out[SD_SRC_RAYLEIGH]) cd.src_out[SD_SRC_RAYLEIGH][o] = v;
            tot += v;
        }
        if (cd.electron) {
            double v = cd.electron[d];
            if (cd.src_out[SD_SRC_ELECTRON]) cd.src_out[SD_SRC_ELECTRON][o] = v;
            tot += v;
        }
        if (cd.line0) tot += cd.line0[o];
        if (cd.line1) tot += cd.line1[o];
        total[o] = tot;
    }
}

}  // namespace

int sd_k3_continuum(sd_ctx *c, const sd_continuum *desc, uint32_t store_mask) {
    const int D = c->D;
    const int64_t W = c->W();
    // pack every small host/device array of the descriptor into one device buffer
    struct Item { const void *src; size_t bytes; size_t off; };
    std::vector<Item> items;
    size_t cursor = 0;
    auto add = [&](const void *p, size_t bytes) -> size_t {
        if (!p) return (size_t)-1;
        size_t off = cursor;
        items.push_back({p, bytes, off});
        cursor += (bytes + 15) & ~(size_t)15;
        return off;
    };
    SD_CHECK(c, desc->n_bf_levels >= 0, SD_ERR_ARG, "negative bf level count");
    SD_CHECK(c, desc->n_bf_levels == 0 || (desc->bf_nu_cut && desc->bf_prefix), SD_ERR_ARG, "bf arrays missing");
    SD_CHECK(c, !desc->ray_c4 == !desc->ray_c6 && !desc->ray_c4 == !desc->ray_c8, SD_ERR_ARG, "Rayleigh needs c4, c6, c8");
    size_t o_cut = add(desc->n_bf_levels ? desc->bf_nu_cut : nullptr, sizeof(double) * desc->n_bf_levels);
    size_t o_pre = add(desc->bf_prefix, sizeof(double) * (desc->n_bf_levels + 1) * D);
    size_t o_ff = add(desc->ff_coef, sizeof(double) * D);
    size_t o_c4 = add(desc->ray_c4, sizeof(double) * D), o_c6 = add(desc->ray_c6, sizeof(double) * D),
           o_c8 = add(desc->ray_c8, sizeof(double) * D);
    size_t o_el = add(desc->electron, sizeof(double) * D);
    size_t o_t[SD_MAX_TABLES][6];
    for (int t = 0; t < desc->n_tables; t++) {
        const sd_table &tb = desc->tables[t];
        SD_CHECK(c, (tb.kind == 1 && tb.nx >= 2 && tb.x && tb.values && tb.depth_scale) ||
                        (tb.kind == 2 && tb.nx >= 2 && tb.ny >= 2 && tb.x && tb.y && tb.values && tb.diag && tb.depth_y &&
                         tb.depth_scale),
                 SD_ERR_ARG, "table %d is malformed", t);
        int ny = tb.kind == 1 ? 1 : tb.ny;
        o_t[t][0] = add(tb.x, sizeof(double) * tb.nx);
        o_t[t][1] = add(tb.kind == 2 ? tb.y : nullptr, sizeof(double) * ny);
        o_t[t][2] = add(tb.values, sizeof(double) * tb.nx * ny);
        o_t[t][3] = add(tb.kind == 2 ? tb.diag : nullptr, (size_t)(tb.nx - 1) * (ny > 1 ? ny - 1 : 1));
        o_t[t][4] = add(tb.kind == 2 ? tb.depth_y : nullptr, sizeof(double) * D);
        o_t[t][5] = add(tb.depth_scale, sizeof(double) * D);
    }
    SD_TRY(sd_ensure(c, c->cont_small, cursor));
    char *base = c->cont_small.as<char>();
    for (const Item &it : items)
        SD_CUDA(c, cudaMemcpyAsync(base + it.off, it.src, it.bytes, cudaMemcpyDefault, c->stream));
    auto dp = [&](size_t off) -> const double * { return off == (size_t)-1 ? nullptr : (const double *)(base + off); };

    ContDev cd{};
    cd.n_bf = desc->n_bf_levels;
    cd.bf_nu_cut = dp(o_cut);
    cd.bf_prefix = dp(o_pre);
    cd.ff_coef = dp(o_ff);
    cd.c4 = dp(o_c4); cd.c6 = dp(o_c6); cd.c8 = dp(o_c8);
    cd.electron = dp(o_el);
    cd.n_tables = desc->n_tables;
    for (int t = 0; t < desc->n_tables; t++) {
        const sd_table &tb = desc->tables[t];
        cd.tab[t].kind = tb.kind; cd.tab[t].nx = tb.nx; cd.tab[t].ny = tb.kind == 1 ? 1 : tb.ny;
        cd.tab[t].x = dp(o_t[t][0]); cd.tab[t].y = dp(o_t[t][1]); cd.tab[t].values = dp(o_t[t][2]);
        cd.tab[t].diag = (const uint8_t *)dp(o_t[t][3]);
        cd.tab[t].depth_y = dp(o_t[t][4]); cd.tab[t].depth_scale = dp(o_t[t][5]);
    }
    bool present[SD_MAX_SOURCES] = {};
    present[SD_SRC_BF] = cd.bf_prefix != nullptr;
    present[SD_SRC_FF] = cd.ff_coef != nullptr;
    present[SD_SRC_RAYLEIGH] = cd.c4 != nullptr;
    present[SD_SRC_ELECTRON] = cd.electron != nullptr;
    for (int t = 0; t < desc->n_tables; t++) present[SD_SRC_TABLE0 + t] = true;
    for (int s = 0; s < SD_MAX_SOURCES; s++) {
        c->have_src[s] = false;
        cd.src_out[s] = nullptr;
        if (present[s] && ((store_mask >> s) & 1u)) {
            SD_TRY(sd_ensure(c, c->src[s], sizeof(double) * D * W));
            cd.src_out[s] = c->src[s].as<double>();
            c->have_src[s] = true;
        }
    }
    cd.line0 = c->have_alpha[0] ? c->alpha_line[0].as<double>() : nullptr;
    cd.line1 = c->have_alpha[1] ? c->alpha_line[1].as<double>() : nullptr;
    SD_TRY(sd_ensure(c, c->total, sizeof(double) * D * W));
    k_continuum<<<(unsigned)((W + 255) / 256), 256, 0, c->stream>>>(D, c->N, c->p0, c->p1, c->nus.as<double>(), cd,
                                                                   c->total.as<double>());
    return sd_launch_check(c, "k_continuum");
}
