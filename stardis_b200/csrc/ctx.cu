// ctx.cu -- context, memory, transfers, result access, elementwise kernels and measurement helpers of the
// C ABI declared in include/stardis_b200.h.
#include <stdarg.h>

#include "sd_internal.h"
#include "sd_math.cuh"

int sd_fail(sd_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

int sd_ensure(sd_ctx *c, DevBuf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return SD_OK;
    if (b.p) {
        SD_CUDA(c, cudaStreamSynchronize(c->stream));
        SD_CUDA(c, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return sd_fail(c, SD_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SD_OK;
}

int sd_upload(sd_ctx *c, DevBuf &b, const void *src, size_t bytes) {
    SD_TRY(sd_ensure(c, b, bytes));
    if (bytes) SD_CUDA(c, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyDefault, c->stream));
    return SD_OK;
}

int sd_launch_check(sd_ctx *c, const char *what) {
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sd_fail(c, SD_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SD_OK;
}

static void free_buf(DevBuf &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

extern "C" {

const char *sd_version(void) { return "stardis_b200 0.1 (sm_100a)"; }

int sd_create(sd_ctx **out, int device) {
    if (!out) return SD_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SD_ERR_CUDA;
    if (device < 0 || device >= n) return SD_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return SD_ERR_CUDA;
    sd_ctx *c = new sd_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return SD_ERR_CUDA;
    }
    c->stream = c->own_stream;
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    for (int k = 0; k < SD_N_PHASES; k++) {
        cudaEventCreate(&c->ph_ev[k][0]);
        cudaEventCreate(&c->ph_ev[k][1]);
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS) != SD_OK ||
        cudaMemset(c->stats.p, 0, sizeof(unsigned long long) * SD_N_STATS) != cudaSuccess) {
        sd_destroy(c);
        return SD_ERR_NOMEM;
    }
    *out = c;
    return SD_OK;
}

void sd_destroy(sd_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf *all[] = {&c->T, &c->ne, &c->nH, &c->nus, &c->d_nu, &c->l_nu, &c->l_Z, &c->l_ion, &c->l_eion, &c->l_eup,
                     &c->l_elo, &c->l_A, &c->l_mass, &c->l_stark, &c->l_waals, &c->l_alpha, &c->gammas, &c->dws,
                     &c->vald_stage, &c->line_idx, &c->rec, &c->win, &c->win_cls, &c->cls_list, &c->cls_off,
                     &c->chunk_cnt, &c->stats, &c->lev_info, &c->fc_tab, &c->edge_tab, &c->edge_keys, &c->edge_unsorted, &c->edge_off, &c->edge_count,
                     &c->line_pre, &c->depth_pre, &c->edge_sort_tmp, &c->alpha_line[0], &c->alpha_line[1], &c->total, &c->cont_small,
                     &c->F, &c->I_nus, &c->ray_small};
    for (DevBuf *b : all) free_buf(*b);
    for (int i = 0; i < SD_MAX_SOURCES; i++) free_buf(c->src[i]);
    for (int k = 0; k < SD_FAR_LEVELS; k++) {
        free_buf(c->tile_geom[k]);
        free_buf(c->far_coef[k]);
        free_buf(c->far_mom[k]);
        for (int h = 0; h < SD_FAR_LEVELS; h++) free_buf(c->far_bkt[k][h]);
    }
    free_buf(c->far_part);
    if (c->h_edge_count) cudaFreeHost(c->h_edge_count);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    for (int k = 0; k < SD_N_PHASES; k++) {
        cudaEventDestroy(c->ph_ev[k][0]);
        cudaEventDestroy(c->ph_ev[k][1]);
    }
    cudaStreamDestroy(c->own_stream);
    delete c;
}

const char *sd_last_error(const sd_ctx *c) { return c ? c->err.c_str() : "null context"; }

int sd_set_stream(sd_ctx *c, void *s) {
    if (!c) return SD_ERR_ARG;
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return SD_OK;
}

int sd_synchronize(sd_ctx *c) {
    if (!c) return SD_ERR_ARG;
    SD_CUDA(c, cudaStreamSynchronize(c->stream));
    return SD_OK;
}

int sd_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return SD_ERR_ARG;
    return cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? SD_OK : SD_ERR_NOMEM;
}
int sd_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? SD_OK : SD_ERR_CUDA; }

int sd_set_atmosphere(sd_ctx *c, int32_t D, const double *T, const double *ne, const double *nH, double vmic) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, D >= 2 && T, SD_ERR_ARG, "sd_set_atmosphere: need n_depth >= 2 and temperatures");
    SD_CUDA(c, cudaSetDevice(c->device));
    if (D != c->D) {
        c->records_ready = false;
        c->have_alpha[0] = c->have_alpha[1] = c->have_total = c->have_F = c->have_broadening = false;
    }
    c->D = D;
    c->vmic = vmic;
    SD_TRY(sd_upload(c, c->T, T, sizeof(double) * D));
    if (ne) SD_TRY(sd_upload(c, c->ne, ne, sizeof(double) * D));
    if (nH) SD_TRY(sd_upload(c, c->nH, nH, sizeof(double) * D));
    if (!ne) {
        SD_TRY(sd_ensure(c, c->ne, sizeof(double) * D));
        SD_CUDA(c, cudaMemsetAsync(c->ne.p, 0, sizeof(double) * D, c->stream));
    }
    if (!nH) {
        SD_TRY(sd_ensure(c, c->nH, sizeof(double) * D));
        SD_CUDA(c, cudaMemsetAsync(c->nH.p, 0, sizeof(double) * D, c->stream));
    }
    return SD_OK;
}

int sd_set_grid(sd_ctx *c, int64_t N, const double *nus, int64_t p0, int64_t p1) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, N >= 2 && N < (int64_t)2147483000 && nus, SD_ERR_ARG, "sd_set_grid: need 2 <= N < 2^31");
    SD_CHECK(c, 0 <= p0 && p0 < p1 && p1 <= N, SD_ERR_ARG, "sd_set_grid: need 0 <= p0 < p1 <= N");
    SD_CUDA(c, cudaSetDevice(c->device));
    c->N = N;
    c->p0 = p0;
    c->p1 = p1;
    c->records_ready = false;
    c->have_alpha[0] = c->have_alpha[1] = c->have_total = c->have_F = false;
    for (int i = 0; i < SD_MAX_SOURCES; i++) c->have_src[i] = false;
    SD_TRY(sd_upload(c, c->nus, nus, sizeof(double) * N));
    return SD_OK;
}

int sd_set_lines(sd_ctx *c, const sd_lines *ln) {
    if (!c || !ln) return SD_ERR_ARG;
    SD_CHECK(c, c->D > 0, SD_ERR_STATE, "sd_set_lines: call sd_set_atmosphere first");
    int64_t L = ln->n_lines;
    SD_CHECK(c, L >= 0 && L * (int64_t)c->D < (int64_t)2147483647, SD_ERR_ARG,
             "sd_set_lines: bad line count (lines x depth points must stay below 2^31: 32-bit pair indices)");
    SD_CHECK(c, L == 0 || ln->nu, SD_ERR_ARG, "sd_set_lines: nu is required");
    SD_CUDA(c, cudaSetDevice(c->device));
    c->L = L;
    c->records_ready = false;
    c->have_broadening = false;
    size_t d8 = sizeof(double) * L;
    SD_TRY(sd_upload(c, c->l_nu, ln->nu, d8));
    if (ln->alpha_line) SD_TRY(sd_upload(c, c->l_alpha, ln->alpha_line, d8 * c->D));
    else SD_TRY(sd_ensure(c, c->l_alpha, d8 * c->D));  // filled by sd_calc_alpha_line_vald
    c->have_alpha_line = ln->alpha_line != nullptr || L == 0;
    if (ln->mass) SD_TRY(sd_upload(c, c->l_mass, ln->mass, d8));
    c->has_atomic_cols = ln->atomic_number && ln->ion_number && ln->ionization_energy && ln->level_energy_upper &&
                         ln->level_energy_lower && ln->A_ul && ln->mass;
    if (c->has_atomic_cols) {
        SD_TRY(sd_upload(c, c->l_Z, ln->atomic_number, sizeof(int64_t) * L));
        SD_TRY(sd_upload(c, c->l_ion, ln->ion_number, sizeof(int64_t) * L));
        SD_TRY(sd_upload(c, c->l_eion, ln->ionization_energy, d8));
        SD_TRY(sd_upload(c, c->l_eup, ln->level_energy_upper, d8));
        SD_TRY(sd_upload(c, c->l_elo, ln->level_energy_lower, d8));
        SD_TRY(sd_upload(c, c->l_A, ln->A_ul, d8));
    }
    c->has_vald_cols = ln->stark && ln->waals;
    if (c->has_vald_cols) {
        SD_TRY(sd_upload(c, c->l_stark, ln->stark, d8));
        SD_TRY(sd_upload(c, c->l_waals, ln->waals, d8));
    }
    return SD_OK;
}

namespace {
constexpr double ALPHA_COEFFICIENT = (sdm::PI * sdm::E_ESU * sdm::E_ESU) / (9.1093837015e-28 * sdm::C_CGS);  // plasma/base.py:35

// count of non-finite line strengths (the reference raises ValueError for them, plasma/base.py:161-164, 293-296)
__device__ __forceinline__ void flag_nonfinite(double a, unsigned long long *__restrict__ stats) {
    const unsigned m = __ballot_sync(__activemask(), !(fabs(a) < INFINITY));
    if (m && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd(&stats[11], (unsigned long long)__popc(m));
}

// one thread per (line, depth), depth fastest: coalesced (L, D) store, broadcast-friendly per-line loads
__global__ void __launch_bounds__(256) k_alpha_line_vald(int64_t L, int D, const double *__restrict__ T,
                                                         const double *__restrict__ line_nu, const double *__restrict__ n_over_u,
                                                         const int64_t *__restrict__ ion_row, const double *__restrict__ gf,
                                                         const double *__restrict__ g_lo, const double *__restrict__ e_low,
                                                         double *__restrict__ alpha, unsigned long long *__restrict__ stats) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= L * D) return;
    const int64_t l = g / D;
    const int d = (int)(g - l * D);
    const double Td = T[d];
    // plasma/base.py:242-247: exp(outer(-E_low, 1 / (T k_B)))
    const double boltz = exp((-e_low[l]) * (1.0 / (Td * sdm::KB_CGS)));
    double n_lower = boltz * n_over_u[ion_row[l] * D + d];   // :256-262 (x g_lo for the long lists)
    if (g_lo) n_lower *= g_lo[l];
    // :272-281: 1 - exp((-h / k_B) * outer(nu, 1 / T))
    const double emis = 1.0 - exp((-sdm::H_CGS / sdm::KB_CGS) * (line_nu[l] * (1.0 / Td)));
    const double a = ALPHA_COEFFICIENT * n_lower * gf[l] * emis;  // :283-291, left to right
    alpha[g] = a;
    flag_nonfinite(a, stats);
}

// AlphaLine (plasma/base.py:130-175): alpha = ALPHA_COEFFICIENT * n_lower * stimulated_emission_factor * f_lu with
// tardis' StimulatedEmissionFactor (third-party tardis release-2024.08.25, plasma/properties/radiative_properties.py;
// source absent offline, restated from its published algorithm):
//   sef = 1 - (g_lower n_upper) / (g_upper n_lower);  0 where n_lower == 0, where it is -inf, and where it is negative
//   for a line whose upper level is metastable.
__global__ void __launch_bounds__(256) k_alpha_line_levels(int64_t L, int D, const double *__restrict__ n_level,
                                                           const double *__restrict__ g_level, const int64_t *__restrict__ lower,
                                                           const int64_t *__restrict__ upper,
                                                           const int64_t *__restrict__ metastable_upper,
                                                           const double *__restrict__ f_lu, double *__restrict__ alpha,
                                                           unsigned long long *__restrict__ stats) {
    const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= L * D) return;
    const int64_t l = g / D;
    const int d = (int)(g - l * D);
    const int64_t lo = lower[l], up = upper[l];
    const double n_lower = n_level[lo * D + d], n_upper = n_level[up * D + d];
    double sef = 1.0 - ((g_level[lo] * n_upper) / (g_level[up] * n_lower));
    if (n_lower == 0.0) sef = 0.0;
    if (sef == -INFINITY) sef = 0.0;
    if (metastable_upper && metastable_upper[l] != 0 && sef < 0.0) sef = 0.0;
    const double a = ALPHA_COEFFICIENT * n_lower * sef * f_lu[l];
    alpha[g] = a;
    flag_nonfinite(a, stats);
}

int strength_prologue(sd_ctx *c, const char *who) {
    SD_CHECK(c, c->D > 0, SD_ERR_STATE, "%s: no atmosphere", who);
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS));
    SD_CUDA(c, cudaMemsetAsync(c->stats.as<unsigned long long>() + 11, 0, sizeof(unsigned long long), c->stream));
    return SD_OK;
}
}  // namespace

int sd_calc_alpha_line_vald(sd_ctx *c, int64_t n_ions, const double *n_over_u, const int64_t *ion_row, const double *gf,
                            const double *g_lo, const double *e_low_erg) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, n_ions > 0 && n_over_u && ion_row && gf, SD_ERR_ARG, "sd_calc_alpha_line_vald: missing input");
    SD_CHECK(c, e_low_erg || c->has_atomic_cols, SD_ERR_ARG,
             "sd_calc_alpha_line_vald: e_low_erg == NULL needs the line table's level_energy_lower column");
    SD_TRY(strength_prologue(c, "sd_calc_alpha_line_vald"));
    const int64_t L = c->L, n = L * c->D;
    if (n == 0) return SD_OK;
    // staged inputs: [n_over_u | ion_row | gf | e_low | g_lo]
    const size_t b_tab = sizeof(double) * n_ions * c->D, b8 = sizeof(double) * L;
    SD_TRY(sd_ensure(c, c->vald_stage, b_tab + 4 * b8));
    char *base = c->vald_stage.as<char>();
    SD_CUDA(c, cudaMemcpyAsync(base, n_over_u, b_tab, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(base + b_tab, ion_row, b8, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(base + b_tab + b8, gf, b8, cudaMemcpyDefault, c->stream));
    if (e_low_erg) SD_CUDA(c, cudaMemcpyAsync(base + b_tab + 2 * b8, e_low_erg, b8, cudaMemcpyDefault, c->stream));
    if (g_lo) SD_CUDA(c, cudaMemcpyAsync(base + b_tab + 3 * b8, g_lo, b8, cudaMemcpyDefault, c->stream));
    sd_phase_begin(c, SD_PH_STRENGTH);
    k_alpha_line_vald<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        L, c->D, c->T.as<double>(), c->l_nu.as<double>(), reinterpret_cast<const double *>(base),
        reinterpret_cast<const int64_t *>(base + b_tab), reinterpret_cast<const double *>(base + b_tab + b8),
        g_lo ? reinterpret_cast<const double *>(base + b_tab + 3 * b8) : nullptr,
        e_low_erg ? reinterpret_cast<const double *>(base + b_tab + 2 * b8) : c->l_elo.as<double>(), c->l_alpha.as<double>(),
        c->stats.as<unsigned long long>());
    SD_TRY(sd_launch_check(c, "k_alpha_line_vald"));
    sd_phase_end(c, SD_PH_STRENGTH);
    c->have_alpha_line = true;
    c->records_ready = false;
    return SD_OK;
}

int sd_calc_alpha_line_levels(sd_ctx *c, int64_t n_levels, const double *level_number_density, const double *g,
                              const int64_t *lower_level_index, const int64_t *upper_level_index,
                              const int64_t *metastable_upper, const double *f_lu) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, n_levels > 0 && level_number_density && g && lower_level_index && upper_level_index && f_lu, SD_ERR_ARG,
             "sd_calc_alpha_line_levels: missing input");
    SD_TRY(strength_prologue(c, "sd_calc_alpha_line_levels"));
    const int64_t L = c->L, n = L * c->D;
    if (n == 0) return SD_OK;
    // staged inputs: [level densities | g | lower | upper | f_lu | metastable]
    const size_t b_tab = sizeof(double) * n_levels * c->D, b_g = sizeof(double) * n_levels, b8 = sizeof(double) * L;
    SD_TRY(sd_ensure(c, c->vald_stage, b_tab + b_g + 4 * b8));
    char *base = c->vald_stage.as<char>();
    SD_CUDA(c, cudaMemcpyAsync(base, level_number_density, b_tab, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(base + b_tab, g, b_g, cudaMemcpyDefault, c->stream));
    char *per_line = base + b_tab + b_g;
    SD_CUDA(c, cudaMemcpyAsync(per_line, lower_level_index, b8, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(per_line + b8, upper_level_index, b8, cudaMemcpyDefault, c->stream));
    SD_CUDA(c, cudaMemcpyAsync(per_line + 2 * b8, f_lu, b8, cudaMemcpyDefault, c->stream));
    if (metastable_upper) SD_CUDA(c, cudaMemcpyAsync(per_line + 3 * b8, metastable_upper, b8, cudaMemcpyDefault, c->stream));
    sd_phase_begin(c, SD_PH_STRENGTH);
    k_alpha_line_levels<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        L, c->D, reinterpret_cast<const double *>(base), reinterpret_cast<const double *>(base + b_tab),
        reinterpret_cast<const int64_t *>(per_line), reinterpret_cast<const int64_t *>(per_line + b8),
        metastable_upper ? reinterpret_cast<const int64_t *>(per_line + 3 * b8) : nullptr,
        reinterpret_cast<const double *>(per_line + 2 * b8), c->l_alpha.as<double>(), c->stats.as<unsigned long long>());
    SD_TRY(sd_launch_check(c, "k_alpha_line_levels"));
    sd_phase_end(c, SD_PH_STRENGTH);
    c->have_alpha_line = true;
    c->records_ready = false;
    return SD_OK;
}

int sd_calc_broadening(sd_ctx *c, uint32_t flags) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, c->D > 0, SD_ERR_STATE, "sd_calc_broadening: no atmosphere");
    SD_CHECK(c, c->has_atomic_cols, SD_ERR_STATE, "sd_calc_broadening: line table lacks the level/ion columns");
    SD_CHECK(c, !(flags & SD_VALD) || c->has_vald_cols, SD_ERR_STATE, "sd_calc_broadening: SD_VALD needs stark/waals");
    SD_CUDA(c, cudaSetDevice(c->device));
    sd_phase_begin(c, SD_PH_K1);
    SD_TRY(sd_k1_broadening(c, flags));
    sd_phase_end(c, SD_PH_K1);
    c->gamma_cols = c->D;
    c->have_broadening = true;
    c->records_ready = false;
    return SD_OK;
}

int sd_set_broadening(sd_ctx *c, const double *gammas, int32_t gamma_cols, const double *dws) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, gammas && dws && (gamma_cols == 1 || gamma_cols == c->D), SD_ERR_ARG, "sd_set_broadening: bad arguments");
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_upload(c, c->gammas, gammas, sizeof(double) * c->L * gamma_cols));
    SD_TRY(sd_upload(c, c->dws, dws, sizeof(double) * c->L * c->D));
    c->gamma_cols = gamma_cols;
    c->have_broadening = true;
    c->records_ready = false;
    return SD_OK;
}

int sd_calc_alpha_line(sd_ctx *c, int32_t slot) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, slot == 0 || slot == 1, SD_ERR_ARG, "sd_calc_alpha_line: slot must be 0 or 1");
    SD_CHECK(c, c->N > 0 && c->D > 0, SD_ERR_STATE, "sd_calc_alpha_line: grid/atmosphere not set");
    SD_CHECK(c, c->have_broadening || c->L == 0, SD_ERR_STATE, "sd_calc_alpha_line: no broadening (K1) yet");
    SD_CHECK(c, c->have_alpha_line || c->L == 0, SD_ERR_STATE,
             "sd_calc_alpha_line: the line table has no alpha_line (pass it to sd_set_lines or call sd_calc_alpha_line_vald)");
    SD_CUDA(c, cudaSetDevice(c->device));
    if (!c->records_ready) SD_TRY(sd_k2_prepare(c));
    SD_TRY(sd_k2_lines(c, slot));
    c->have_alpha[slot] = true;
    return SD_OK;
}

int sd_set_farfield(sd_ctx *c, int32_t on) {
    if (!c) return SD_ERR_ARG;
    if (c->farfield != (on != 0)) c->records_ready = false;  // the far-capable classes and edge lists belong to the preparation pass
    c->farfield = on != 0;
    return SD_OK;
}

int sd_set_line_stats(sd_ctx *c, int32_t on) {
    if (!c) return SD_ERR_ARG;
    c->line_stats = on != 0;
    return SD_OK;
}

int sd_line_stats_ex(sd_ctx *c, int64_t out[16]) {
    if (!c || !out) return SD_ERR_ARG;
    SD_CHECK(c, c->stats.p, SD_ERR_STATE, "sd_line_stats: nothing computed yet");
    unsigned long long h[SD_N_STATS];
    SD_CUDA(c, cudaMemcpyAsync(h, c->stats.p, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    SD_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < SD_N_STATS; i++) out[i] = (int64_t)h[i];
    return SD_OK;
}

int sd_line_stats(sd_ctx *c, int64_t out[8]) {
    int64_t h[SD_N_STATS];
    SD_TRY(sd_line_stats_ex(c, h));
    for (int i = 0; i < 8; i++) out[i] = h[i];
    out[0] += h[8];  // region-I evaluations the far-field expansion stands for: counted as the reference would
    return SD_OK;
}

int sd_phase_times(sd_ctx *c, float out_ms[8]) {
    if (!c || !out_ms) return SD_ERR_ARG;
    SD_CUDA(c, cudaSetDevice(c->device));
    for (int k = 0; k < SD_N_PHASES; k++) {
        out_ms[k] = -1.0f;
        if (!c->ph_rec[k]) continue;
        SD_CUDA(c, cudaEventSynchronize(c->ph_ev[k][1]));
        SD_CUDA(c, cudaEventElapsedTime(&out_ms[k], c->ph_ev[k][0], c->ph_ev[k][1]));
    }
    return SD_OK;
}

int sd_calc_continuum(sd_ctx *c, const sd_continuum *desc, uint32_t store_mask) {
    if (!c || !desc) return SD_ERR_ARG;
    SD_CHECK(c, c->N > 0 && c->D > 0, SD_ERR_STATE, "sd_calc_continuum: grid/atmosphere not set");
    SD_CHECK(c, desc->n_tables >= 0 && desc->n_tables <= SD_MAX_TABLES, SD_ERR_ARG, "sd_calc_continuum: too many tables");
    SD_CUDA(c, cudaSetDevice(c->device));
    sd_phase_begin(c, SD_PH_K3);
    SD_TRY(sd_k3_continuum(c, desc, store_mask));
    sd_phase_end(c, SD_PH_K3);
    c->have_total = true;
    return SD_OK;
}

int sd_raytrace(sd_ctx *c, int32_t n_theta, const double *ray_ds, const double *weights, int32_t inward, double scale,
                int32_t track) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, n_theta >= 1 && ray_ds && weights, SD_ERR_ARG, "sd_raytrace: bad arguments");
    SD_CHECK(c, c->have_total, SD_ERR_STATE, "sd_raytrace: no total opacity (sd_calc_continuum / sd_set_total)");
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_k4_raytrace(c, n_theta, ray_ds, weights, inward, scale, track));
    sd_phase_end(c, SD_PH_K4);
    c->have_F = true;
    c->n_theta_tracked = track ? n_theta : 0;
    return SD_OK;
}

static int find_buffer(sd_ctx *c, int which, DevBuf **b, int64_t *rows, int64_t *cols) {
    int64_t W = c->W();
    switch (which) {
        case SD_BUF_GAMMAS:
            SD_CHECK(c, c->have_broadening, SD_ERR_STATE, "gammas not computed");
            *b = &c->gammas; *rows = c->L; *cols = c->gamma_cols; return SD_OK;
        case SD_BUF_DOPPLER:
            SD_CHECK(c, c->have_broadening, SD_ERR_STATE, "doppler widths not computed");
            *b = &c->dws; *rows = c->L; *cols = c->D; return SD_OK;
        case SD_BUF_ALPHA_LINE:
        case SD_BUF_ALPHA_MOLECULE: {
            int s = which - SD_BUF_ALPHA_LINE;
            SD_CHECK(c, c->have_alpha[s], SD_ERR_STATE, "alpha_line slot %d not computed", s);
            *b = &c->alpha_line[s]; *rows = c->D; *cols = W; return SD_OK;
        }
        case SD_BUF_LINE_STRENGTH:
            SD_CHECK(c, c->have_alpha_line, SD_ERR_STATE, "line strengths not set");
            *b = &c->l_alpha; *rows = c->L; *cols = c->D; return SD_OK;
        case SD_BUF_NUS:
            SD_CHECK(c, c->N > 0, SD_ERR_STATE, "no grid");
            *b = &c->nus; *rows = 1; *cols = c->N; return SD_OK;
        case SD_BUF_TOTAL:
            SD_CHECK(c, c->have_total, SD_ERR_STATE, "total opacity not computed");
            *b = &c->total; *rows = c->D; *cols = W; return SD_OK;
        case SD_BUF_F_NU:
            SD_CHECK(c, c->have_F, SD_ERR_STATE, "F_nu not computed");
            *b = &c->F; *rows = c->D; *cols = W; return SD_OK;
        case SD_BUF_I_NUS:
            SD_CHECK(c, c->have_F && c->n_theta_tracked > 0, SD_ERR_STATE, "I_nus not tracked");
            *b = &c->I_nus; *rows = c->D; *cols = W * c->n_theta_tracked; return SD_OK;
        default:
            if (which >= SD_BUF_SOURCE0 && which < SD_BUF_SOURCE0 + SD_MAX_SOURCES) {
                int s = which - SD_BUF_SOURCE0;
                SD_CHECK(c, c->have_src[s], SD_ERR_STATE, "continuum source %d was not stored", s);
                *b = &c->src[s]; *rows = c->D; *cols = W; return SD_OK;
            }
    }
    return sd_fail(c, SD_ERR_ARG, "unknown buffer id %d", which);
}

int sd_get(sd_ctx *c, int32_t which, double *dst, int64_t count) {
    if (!c || !dst) return SD_ERR_ARG;
    DevBuf *b; int64_t rows, cols;
    SD_TRY(find_buffer(c, which, &b, &rows, &cols));
    SD_CHECK(c, count == rows * cols, SD_ERR_ARG, "sd_get: count %lld != buffer size %lld", (long long)count, (long long)(rows * cols));
    SD_CUDA(c, cudaSetDevice(c->device));
    if (count) SD_CUDA(c, cudaMemcpyAsync(dst, b->p, sizeof(double) * count, cudaMemcpyDefault, c->stream));
    return SD_OK;
}

int sd_get_row(sd_ctx *c, int32_t which, int32_t row, double *dst, int64_t count) {
    if (!c || !dst) return SD_ERR_ARG;
    DevBuf *b; int64_t rows, cols;
    SD_TRY(find_buffer(c, which, &b, &rows, &cols));
    if (row < 0) row += (int32_t)rows;
    SD_CHECK(c, row >= 0 && row < rows && count == cols, SD_ERR_ARG, "sd_get_row: bad row/count");
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_CUDA(c, cudaMemcpyAsync(dst, b->as<double>() + (size_t)row * cols, sizeof(double) * cols, cudaMemcpyDefault, c->stream));
    return SD_OK;
}

int sd_set_total(sd_ctx *c, const double *total, int64_t count) {
    if (!c || !total) return SD_ERR_ARG;
    SD_CHECK(c, c->N > 0 && c->D > 0 && count == c->D * c->W(), SD_ERR_ARG, "sd_set_total: count must be D*(p1-p0)");
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_upload(c, c->total, total, sizeof(double) * count));
    c->have_total = true;
    return SD_OK;
}

int sd_buffer(sd_ctx *c, int32_t which, void **ptr, int64_t *count) {
    if (!c || !ptr || !count) return SD_ERR_ARG;
    DevBuf *b; int64_t rows, cols;
    SD_TRY(find_buffer(c, which, &b, &rows, &cols));
    *ptr = b->p;
    *count = rows * cols;
    return SD_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------- elementwise kernels
namespace {

// Operands may live on the host: stage them through scratch device buffers.
struct Staged {
    sd_ctx *c;
    DevBuf in[8];
    DevBuf out[4];
    ~Staged() {
        cudaStreamSynchronize(c->stream);
        for (auto &b : in) free_buf(b);
        for (auto &b : out) free_buf(b);
    }
};

__global__ void k_ew_faddeeva(int64_t n, const double *zr, const double *zi, double *wr, double *wi) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) sdm::humlicek_complex(zr[i], zi[i], wr[i], wi[i]);
}
__global__ void k_ew_voigt(int64_t n, const double *dnu, const double *dw, const double *g, double *phi) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) phi[i] = sdm::voigt_profile(dnu[i], dw[i], g[i]);
}
__global__ void k_ew_doppler(int64_t n, const double *nu, const double *T, const double *m, double vmic, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = sdm::doppler_width(nu[i], T[i], m[i], vmic);
}
__global__ void k_ew_neff(int64_t n, const double *z, const double *ei, const double *el, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = sdm::n_effective(z[i], ei[i], el[i]);
}
__global__ void k_ew_lstark(int64_t n, const double *a, const double *b, const double *ne, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = sdm::gamma_linear_stark(a[i], b[i], ne[i]);
}
__global__ void k_ew_qstark(int64_t n, const double *z, const double *a, const double *b, const double *ne, const double *T, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = sdm::gamma_quadratic_stark(z[i], a[i], b[i], ne[i], T[i]);
}
__global__ void k_ew_vdw(int64_t n, const double *z, const double *a, const double *b, const double *T, const double *nH, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) o[i] = sdm::gamma_van_der_waals(z[i], a[i], b[i], T[i], nH[i]);
}
__global__ void k_ew_blackbody(int D, int64_t N, const double *nus, const double *T, double *o) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int d = blockIdx.y;
    if (i < N) o[(int64_t)d * N + i] = sdm::planck(nus[i], T[d]);
}
__global__ void k_ew_weights(int64_t n, const double *tau, double *w0, double *w1, double *w2) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) sdm::rt_weights(tau[i], w0[i], w1[i], w2[i]);
}

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

#define EW_BEGIN(nin, nout)                                        \
    if (!c) return SD_ERR_ARG;                                     \
    SD_CHECK(c, n >= 0, SD_ERR_ARG, "negative element count");      \
    SD_CUDA(c, cudaSetDevice(c->device));                          \
    if (n == 0) return SD_OK;                                      \
    Staged st{c};                                                  \
    const void *ins[] = {EW_INS};                                  \
    void *outs[] = {EW_OUTS};                                      \
    for (int k = 0; k < nin; k++) SD_TRY(sd_upload(c, st.in[k], ins[k], sizeof(double) * n)); \
    for (int k = 0; k < nout; k++) SD_TRY(sd_ensure(c, st.out[k], sizeof(double) * n));
#define EW_END(nout, what)                                         \
    SD_TRY(sd_launch_check(c, what));                              \
    for (int k = 0; k < nout; k++)                                 \
        SD_CUDA(c, cudaMemcpyAsync(outs[k], st.out[k].p, sizeof(double) * n, cudaMemcpyDefault, c->stream)); \
    return SD_OK;
#define I(k) st.in[k].as<double>()
#define O(k) st.out[k].as<double>()

extern "C" {

int sd_ew_faddeeva(sd_ctx *c, int64_t n, const double *zr, const double *zi, double *wr, double *wi) {
#define EW_INS zr, zi
#define EW_OUTS wr, wi
    EW_BEGIN(2, 2)
    k_ew_faddeeva<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), O(0), O(1));
    EW_END(2, "k_ew_faddeeva")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_voigt_profile(sd_ctx *c, int64_t n, const double *dnu, const double *dw, const double *g, double *phi) {
#define EW_INS dnu, dw, g
#define EW_OUTS phi
    EW_BEGIN(3, 1)
    k_ew_voigt<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), O(0));
    EW_END(1, "k_ew_voigt")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_doppler_width(sd_ctx *c, int64_t n, const double *nu, const double *T, const double *m, double vmic, double *out) {
#define EW_INS nu, T, m
#define EW_OUTS out
    EW_BEGIN(3, 1)
    k_ew_doppler<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), vmic, O(0));
    EW_END(1, "k_ew_doppler")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_n_effective(sd_ctx *c, int64_t n, const double *z, const double *ei, const double *el, double *out) {
#define EW_INS z, ei, el
#define EW_OUTS out
    EW_BEGIN(3, 1)
    k_ew_neff<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), O(0));
    EW_END(1, "k_ew_neff")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_gamma_linear_stark(sd_ctx *c, int64_t n, const double *a, const double *b, const double *ne, double *out) {
#define EW_INS a, b, ne
#define EW_OUTS out
    EW_BEGIN(3, 1)
    k_ew_lstark<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), O(0));
    EW_END(1, "k_ew_lstark")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_gamma_quadratic_stark(sd_ctx *c, int64_t n, const double *z, const double *a, const double *b, const double *ne,
                                const double *T, double *out) {
#define EW_INS z, a, b, ne, T
#define EW_OUTS out
    EW_BEGIN(5, 1)
    k_ew_qstark<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), I(3), I(4), O(0));
    EW_END(1, "k_ew_qstark")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_gamma_van_der_waals(sd_ctx *c, int64_t n, const double *z, const double *a, const double *b, const double *T,
                              const double *nH, double *out) {
#define EW_INS z, a, b, T, nH
#define EW_OUTS out
    EW_BEGIN(5, 1)
    k_ew_vdw<<<nblk(n), 256, 0, c->stream>>>(n, I(0), I(1), I(2), I(3), I(4), O(0));
    EW_END(1, "k_ew_vdw")
#undef EW_INS
#undef EW_OUTS
}
int sd_ew_calc_weights(sd_ctx *c, int64_t n, const double *tau, double *w0, double *w1, double *w2) {
#define EW_INS tau
#define EW_OUTS w0, w1, w2
    EW_BEGIN(1, 3)
    k_ew_weights<<<nblk(n), 256, 0, c->stream>>>(n, I(0), O(0), O(1), O(2));
    EW_END(3, "k_ew_weights")
#undef EW_INS
#undef EW_OUTS
}

int sd_ew_blackbody(sd_ctx *c, int32_t D, int64_t N, const double *nus, const double *T, double *out) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, D > 0 && N > 0 && D < 65536, SD_ERR_ARG, "sd_ew_blackbody: bad shape");
    SD_CUDA(c, cudaSetDevice(c->device));
    Staged st{c};
    SD_TRY(sd_upload(c, st.in[0], nus, sizeof(double) * N));
    SD_TRY(sd_upload(c, st.in[1], T, sizeof(double) * D));
    SD_TRY(sd_ensure(c, st.out[0], sizeof(double) * N * D));
    k_ew_blackbody<<<dim3(nblk(N), D), 256, 0, c->stream>>>(D, N, I(0), I(1), O(0));
    SD_TRY(sd_launch_check(c, "k_ew_blackbody"));
    SD_CUDA(c, cudaMemcpyAsync(out, st.out[0].p, sizeof(double) * N * D, cudaMemcpyDefault, c->stream));
    return SD_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------- spectrum convolution
namespace {
// out[i] = sum_j w[j] x[reflect(i - j + m/2)], m odd: scipy.ndimage.convolve1d(x, w) with its default mode "reflect"
// (half-sample symmetric: d c b a | a b c d | d c b a), as rotation_broadening calls it (broadening.py:866-874)
__global__ void __launch_bounds__(256) k_convolve1d_reflect(int64_t n, const double *__restrict__ x, int m,
                                                            const double *__restrict__ w, double *__restrict__ out) {
    extern __shared__ double s_w[];
    for (int k = threadIdx.x; k < m; k += blockDim.x) s_w[k] = w[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int h = m / 2;
    double acc = 0.0;
    for (int j = 0; j < m; j++) {
        int64_t k = i - j + h;
        while (k < 0 || k >= n) k = (k < 0) ? -k - 1 : 2 * n - k - 1;
        acc += s_w[j] * x[k];  // scipy accumulates in kernel order as well; agreement ~1e-16 relative
    }
    out[i] = acc;
}
}  // namespace

extern "C" int sd_convolve1d_reflect(sd_ctx *c, int64_t n, const double *x, int32_t m, const double *weights, double *out) {
    if (!c) return SD_ERR_ARG;
    SD_CHECK(c, n > 0 && x && weights && out && m > 0 && (m & 1) && m <= 6000, SD_ERR_ARG,
             "sd_convolve1d_reflect: need n > 0 and an odd kernel length <= 6000");
    SD_CUDA(c, cudaSetDevice(c->device));
    Staged st{c};
    SD_TRY(sd_upload(c, st.in[0], x, sizeof(double) * n));
    SD_TRY(sd_upload(c, st.in[1], weights, sizeof(double) * m));
    SD_TRY(sd_ensure(c, st.out[0], sizeof(double) * n));
    k_convolve1d_reflect<<<nblk(n), 256, sizeof(double) * m, c->stream>>>(n, st.in[0].as<double>(), m, st.in[1].as<double>(),
                                                                          st.out[0].as<double>());
    SD_TRY(sd_launch_check(c, "k_convolve1d_reflect"));
    SD_CUDA(c, cudaMemcpyAsync(out, st.out[0].p, sizeof(double) * n, cudaMemcpyDefault, c->stream));
    return SD_OK;
}

// ------------------------------------------------------------------------------- measurement helpers
namespace {
// 8 independent FMA chains per thread, no memory traffic: the FP64 FMA-pipe ceiling of the chip.
__global__ void __launch_bounds__(256) k_dfma(int iters, double seed, double *sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-7;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

// FP64 issue-rate probes.  mode 1: every DFMA reads three DISTINCT register pairs (no operand reuse);
// mode 2: two distinct pairs + one reused; mode 3: mode 0 plus one MUFU.RCP64H per 8 DFMA.
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64_probe(int iters, double seed, double *sink) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed + threadIdx.x + i;
        b[i] = 0.999999 + 1e-9 * i + 1e-12 * threadIdx.x;
        c[i] = 1e-7 * (i + 1);
    }
    const double m = 0.999999, k = 1e-7;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 1) a[i] = fma(b[i], c[i], a[i]);
                else if (MODE == 2) a[i] = fma(a[i], b[i], k);
                else a[i] = fma(a[i], m, k);
            }
            if (MODE == 3) {
                double r;
                asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a[u]));
                c[u] = r;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + c[i];
    if (s == 123.456) sink[0] = s;
}

// The far-wing evaluation of k_lines in isolation: 8 pixels per thread, entry constants in registers (perturbed per
// iteration so nothing is hoisted), no shared memory, no window logic.  VARIANT 0: full body; 1: without the
// reciprocal (den used directly: no MUFU, no Newton); 2: MUFU seed only (no Newton).
template <int VARIANT>
__global__ void __launch_bounds__(256, 2) k_fareval_probe(int iters, double seed, double *sink) {
    double nu[8], acc[8];
#pragma unroll
    for (int p = 0; p < 8; p++) {
        nu[p] = 5.0e14 + 1.0e9 * (threadIdx.x + 256 * p);
        acc[p] = 0.0;
    }
    double inv_dw = 4.0e-10 * seed, xl = 1.9e5, eb = -0.98, ec = 0.26, Kc = 1e-3, Kf = 2e-3;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int p = 0; p < 8; p++) {
            double x = fma(nu[p], inv_dw, -xl);
            double q = x * x;
            double den = fma(q, q + eb, ec);
            double num = fma(Kf, q, Kc);
            double r;
            if (VARIANT == 1) r = den;
            else if (VARIANT == 2) { asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den)); }
            else r = sdm::rcp_fast2(den);
            acc[p] = fma(num, r, acc[p]);
        }
        xl += 1.0; eb += 1e-9; Kf += 1e-12;  // 3 extra FP64 per 8 evaluations (accounted for below)
    }
    double s = 0;
#pragma unroll
    for (int p = 0; p < 8; p++) s += acc[p];
    if (s == 123.456) sink[0] = s;
}

__global__ void k_debug_rcp(int64_t n, const double *x, double *seed, double *quad, double *cubic) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i]));
    seed[i] = r;
    quad[i] = sdm::rcp_fast2(x[i]);
    cubic[i] = sdm::rcp_fast(x[i]);
}
}  // namespace

extern "C" int sd_debug_rcp(sd_ctx *c, int64_t n, const double *x, double *seed, double *quad, double *cubic) {
#define EW_INS x
#define EW_OUTS seed, quad, cubic
    EW_BEGIN(1, 3)
    k_debug_rcp<<<nblk(n), 256, 0, c->stream>>>(n, I(0), O(0), O(1), O(2));
    EW_END(3, "k_debug_rcp")
#undef EW_INS
#undef EW_OUTS
}

extern "C" {

int sd_bench_dfma(sd_ctx *c, int32_t iters, double *tflops) {
    if (!c || !tflops || iters <= 0) return SD_ERR_ARG;
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS));
    int blocks = c->sm_count * 8;
    k_dfma<<<blocks, 256, 0, c->stream>>>(iters / 8 + 1, 1.0, c->stats.as<double>() + 7);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        SD_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        k_dfma<<<blocks, 256, 0, c->stream>>>(iters, 1.0, c->stats.as<double>() + 7);
        SD_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        SD_CUDA(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        SD_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (ms < best) best = ms;
    }
    SD_TRY(sd_launch_check(c, "k_dfma"));
    double fmas = (double)blocks * 256.0 * (double)iters * 64.0;
    *tflops = 2.0 * fmas / (best * 1e-3) / 1e12;
    return SD_OK;
}

int sd_bench_fp64(sd_ctx *c, int32_t mode, int32_t iters, double *tflops) {
    if (!c || !tflops || iters <= 0 || mode < 0 || mode > 3) return SD_ERR_ARG;
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS));
    int blocks = c->sm_count * 8;
    double *sink = c->stats.as<double>() + 7;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        SD_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        switch (mode) {
            case 0: k_fp64_probe<0><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink); break;
            case 1: k_fp64_probe<1><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink); break;
            case 2: k_fp64_probe<2><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink); break;
            default: k_fp64_probe<3><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink); break;
        }
        SD_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        SD_CUDA(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        SD_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    SD_TRY(sd_launch_check(c, "k_fp64_probe"));
    *tflops = 2.0 * (double)blocks * 256.0 * (double)iters * 64.0 / (best * 1e-3) / 1e12;
    return SD_OK;
}

int sd_bench_fareval(sd_ctx *c, int32_t variant, int32_t iters, double *gevals) {
    if (!c || !gevals || iters <= 0) return SD_ERR_ARG;
    SD_CUDA(c, cudaSetDevice(c->device));
    SD_TRY(sd_ensure(c, c->stats, sizeof(unsigned long long) * SD_N_STATS));
    int blocks = c->sm_count * 2;
    double *sink = c->stats.as<double>() + 7;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        SD_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        if (variant == 1) k_fareval_probe<1><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink);
        else if (variant == 2) k_fareval_probe<2><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink);
        else k_fareval_probe<0><<<blocks, 256, 0, c->stream>>>(iters, 1.0, sink);
        SD_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        SD_CUDA(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        SD_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    SD_TRY(sd_launch_check(c, "k_fareval_probe"));
    *gevals = (double)blocks * 256.0 * (double)iters * 8.0 / (best * 1e-3) / 1e9;
    return SD_OK;
}

int64_t sd_launch_count(const sd_ctx *c) { return c ? c->launches : 0; }

int sd_timer_start(sd_ctx *c) {
    if (!c) return SD_ERR_ARG;
    SD_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return SD_OK;
}
int sd_timer_stop(sd_ctx *c, float *ms) {
    if (!c || !ms) return SD_ERR_ARG;
    SD_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    SD_CUDA(c, cudaEventSynchronize(c->ev1));
    SD_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SD_OK;
}

}  // extern "C"
