/*
 * stardis_b200.h -- C ABI of the B200-native STARDIS hot path (libstardis_b200.so).
 *
 * The reference (tardis-sn/stardis) has NO FFI layer for this path: the boundary is two Python callables,
 *   calc_alphas(stellar_plasma, stellar_model, stellar_radiation_field, opacity_config)
 *       stardis/radiation_field/opacities/opacities_solvers/base.py:630-740
 *   raytrace(stellar_model, stellar_radiation_field)
 *       stardis/radiation_field/radiation_field_solvers/base.py:271-346
 * which call numba-jitted array functions.  Every entry point below names the reference function (file:line,
 * relative to stardis/) whose array-level work it replaces; the Python side (stardis_b200/) keeps the
 * reference signatures and binds this ABI with ctypes (see INTEGRATION.md for the binding a maintainer
 * of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all real data is fp64, integer columns int64, C-contiguous.
 *   - every data pointer may be a HOST pointer (pageable or pinned) or a DEVICE pointer on the context's
 *     GPU: transfers use cudaMemcpyAsync(..., cudaMemcpyDefault) on the context's stream (UVA).
 *   - every function returns SD_OK (0) or a negative error code; sd_last_error() gives the message.
 *   - work is enqueued on ONE stream per context (sd_set_stream lets the caller supply it, e.g. the
 *     current torch stream); no entry point synchronises the host except sd_synchronize() and copies
 *     into pageable host memory.
 *   - one context per GPU / per rank; contexts are not thread-safe.
 *   - depth index 0 is the deepest point (io/model/marcs.py:203-205); frequency grids are DESCENDING
 *     (ascending wavelengths), as the reference's line kernel assumes (opacities_solvers/base.py:522-558).
 */
#ifndef STARDIS_B200_H
#define STARDIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_CUDA (-1)
#define SD_ERR_ARG (-2)
#define SD_ERR_STATE (-3)
#define SD_ERR_NOMEM (-4)

/* broadening flags == membership in config.opacity.line.broadening (broadening.py:688-691) */
#define SD_LINEAR_STARK 1u
#define SD_QUADRATIC_STARK 2u
#define SD_VAN_DER_WAALS 4u
#define SD_RADIATION 8u
#define SD_VALD 16u /* use the VALD stark/waals columns: calc_vald_gamma, broadening.py:1009-1085 */

typedef struct sd_ctx sd_ctx;

/* ---- context ----------------------------------------------------------------------------------- */
int sd_create(sd_ctx **ctx, int device);
void sd_destroy(sd_ctx *ctx);
const char *sd_last_error(const sd_ctx *ctx);
const char *sd_version(void);
int sd_set_stream(sd_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = the context's own stream;
                                                     * pass cudaStreamLegacy (0x1) for the legacy default stream */
int sd_synchronize(sd_ctx *ctx);
/* pinned host staging buffers for callers without torch (cudaHostAlloc / cudaFreeHost) */
int sd_host_alloc(void **ptr, int64_t bytes);
int sd_host_free(void *ptr);

/* ---- inputs ------------------------------------------------------------------------------------ */
/* Per-depth state read by the broadening and formal-solver kernels.
 * T: stellar_model.temperatures; n_e: stellar_plasma.electron_densities; n_HI:
 * stellar_plasma.ion_number_density.loc[1,0]; vmic_cgs: stellar_model.microturbulence.cgs
 * (broadening.py:706-730). */
int sd_set_atmosphere(sd_ctx *ctx, int32_t n_depth, const double *T, const double *n_e, const double *n_HI,
                      double vmic_cgs);

/* Frequency grid.  nus[N] is the GLOBAL descending grid; this context evaluates pixels [p0, p1) only
 * (nu sharding across GPUs).  Window centres, half-widths and d_nu are always global quantities
 * (opacities_solvers/base.py:522-575), so a shard equals the same columns of a single-GPU run. */
int sd_set_grid(sd_ctx *ctx, int64_t N, const double *nus, int64_t p0, int64_t p1);

/* Line table, ascending in nu and already restricted to [min nu, max nu] of the grid
 * (opacities_solvers/base.py:392-421).  alpha_line is (L, D) row-major [cm^-1 Hz].  Columns that a
 * given broadening mode does not need may be NULL (stark/waals without SD_VALD; everything but nu,
 * mass and alpha_line when sd_set_broadening supplies gammas). */
typedef struct sd_lines {
    int64_t n_lines;
    const double *nu;
    const int64_t *atomic_number;
    const int64_t *ion_number;          /* 0 = neutral; kernels use ion_number + 1 (broadening.py:708-709) */
    const double *ionization_energy;    /* erg */
    const double *level_energy_upper;   /* erg */
    const double *level_energy_lower;   /* erg */
    const double *A_ul;
    const double *mass;                 /* g; atomic, or sum of both constituents for molecules (:808-819) */
    const double *stark;                /* VALD log10 Stark parameter or NULL */
    const double *waals;                /* VALD van der Waals code or NULL */
    const double *alpha_line;           /* (L, D); NULL = filled on the device by sd_calc_alpha_line_vald */
} sd_lines;
int sd_set_lines(sd_ctx *ctx, const sd_lines *lines);

/* Line strengths of a VALD linelist on the device -- the producer of `alpha_line` immediately upstream of the path
 * (SURVEY 8f rank 1): stardis/plasma/base.py:178-321 (AlphaLineVald) and :324-455 (AlphaLineShortlistVald),
 *   alpha[l, d] = (pi e^2 / m_e c) * (N_ion / U)[ion_row[l], d] * exp(-E_low[l] / k T_d) * [g_lo] * gf[l]
 *                 * (1 - exp(-h nu_l / k T_d)),
 * in the reference's order of operations.  Needs the atmosphere (T) and the line table (nu; alpha_line may be NULL).
 * n_over_u: (n_ions, D) ion number density / partition function; ion_row[l]: row of that table for line l;
 * gf[l]: f_lu = 10^log_gf / g_lo with g_lo[l] given (long lists) or 10^log_gf with g_lo == NULL (short lists);
 * e_low_erg[l]: lower level energy [erg] (NULL = the line table's level_energy_lower column, which holds the same
 * numbers in the reference's lines_from_linelist).  Replaces the (L, D) host->device upload of alpha_line by O(L)
 * inputs.  Non-finite results are counted (sd_line_stats_ex out[11]); the reference raises ValueError for them. */
int sd_calc_alpha_line_vald(sd_ctx *ctx, int64_t n_ions, const double *n_over_u, const int64_t *ion_row, const double *gf,
                            const double *g_lo, const double *e_low_erg);
/* Line strengths of the tardis line list on the device: stardis/plasma/base.py:130-175 (AlphaLine),
 *   alpha[l, d] = (pi e^2 / m_e c) * n_lower[l, d] * stimulated_emission_factor[l, d] * f_lu[l],
 * n_lower / n_upper = rows lower_level_index[l] / upper_level_index[l] of level_number_density (n_levels, D), and the
 * stimulated emission factor of tardis (release-2024.08.25, plasma/properties/radiative_properties.py):
 * 1 - (g_lower n_upper) / (g_upper n_lower), set to 0 where n_lower == 0, where it is -inf, and where it is negative for
 * a line whose upper level is metastable (metastable_upper[l] != 0; NULL = no metastable levels).  g: [n_levels]. */
int sd_calc_alpha_line_levels(sd_ctx *ctx, int64_t n_levels, const double *level_number_density, const double *g,
                              const int64_t *lower_level_index, const int64_t *upper_level_index,
                              const int64_t *metastable_upper, const double *f_lu);

/* ---- K1: broadening (calc_gamma broadening.py:550-656, calc_vald_gamma :1009-1085,
 *          calc_doppler_width :32-71) -> gammas (L,D), doppler_widths (L,D) on the device ------------- */
int sd_calc_broadening(sd_ctx *ctx, uint32_t flags);
/* Caller-supplied broadening instead (the calc_alan_entries(D, nus, line_nus, doppler_widths, gammas,
 * alphas) calling convention, opacities_solvers/base.py:487-494).  gamma_cols is D or 1
 * (radiation-only molecular branch, :547-551). */
int sd_set_broadening(sd_ctx *ctx, const double *gammas, int32_t gamma_cols, const double *doppler_widths);

/* ---- K2: windowed Voigt accumulation (calc_alan_entries opacities_solvers/base.py:487-627 with
 *          voigt_profile / _faddeeva voigt.py:17-155) -> alpha_line[slot] (D, p1-p0) --------------- */
/* slot 0 = atomic lines ("alpha_line_at_nu"), slot 1 = molecular lines ("molecule_alpha_line_at_nu"). */
int sd_calc_alpha_line(sd_ctx *ctx, int32_t slot);
/* Far-field expansion of region-I wings per pixel tile (default ON).  On a hierarchy of pixel tiles (64 * 8^k
 * pixels, k = 0..3) a (line, depth) pair whose window covers a whole tile and whose line centre lies at least two
 * tiles away is not evaluated pixel by pixel there: its profile (two Lorentzians in Humlicek region I) is analytic
 * over the tile, the degree-31 Taylor polynomials of all such pairs of a tile are summed once -- through multipole
 * moments of the source tiles and tile-to-tile translations (a 1-D fast multipole method) for pairs whose window
 * covers the whole neighbourhood, by direct expansion otherwise -- and the polynomial is added per pixel (relative
 * deviation from the direct evaluation <= 6e-12 per pair in the worst geometry, all terms positive).  Grids whose
 * tile widths do not vary smoothly lose the upper hierarchy levels automatically.  sd_set_farfield(ctx, 0) evaluates
 * every (line, depth, pixel) triple directly, exactly like the reference's loop. */
int sd_set_farfield(sd_ctx *ctx, int32_t on);
/* Statistics.  sd_set_line_stats(ctx, 1) makes the following sd_calc_alpha_line calls run the counting
 * instantiation of the kernel (slower; for tests and workload descriptions only).  sd_line_stats synchronises:
 * out[0..3] = Voigt evaluations per Humlicek region I..IV, out[4] = (line, depth) pairs with a non-empty
 * window, out[5] = pairs wider than the narrow class, out[6] = pairs with a zero Doppler width (the reference
 * raises ZeroDivisionError for those, voigt.py:148), out[7] reserved. */
int sd_set_line_stats(sd_ctx *ctx, int32_t on);
int sd_line_stats(sd_ctx *ctx, int64_t out[8]);
/* EXECUTED work of the last counting pass (bench.py's roofline of the default, far-field mode), raw counters:
 * out[0..3] = Voigt evaluations k_lines performed itself per Humlicek region I..IV (pixels of this context's range),
 * out[4..6] as sd_line_stats, out[8] = region-I evaluations the far-field expansions stand for
 * (sd_line_stats()[0] = out[0] + out[8]), out[9] = direct far-field expansions performed ((pair, tile) products),
 * out[10] = Taylor terms summed over those expansions, out[11] = non-finite line strengths written by the last
 * sd_calc_alpha_line_vald / sd_calc_alpha_line_levels call, out[12] = multipole expansions ((pair, level) products,
 * 32 moments each), out[13] = matrix-row steps of the tile-to-tile translations per lane ((source, target, depth, k)
 * products: one FMA on each of 32 lanes), out[7], out[14..15] reserved. */
int sd_line_stats_ex(sd_ctx *ctx, int64_t out[16]);

/* ---- K3: continuum terms fused in one depth x nu pass + total ---------------------------------------- */
/* 1-D or 2-D cross-section table (sigma_file, opacities_solvers/util.py:14-108).
 * kind 1: np.interp over x (clamped to the end values), value * depth_scale[d].
 * kind 2: piecewise-linear interpolation on the Delaunay split of the rectangular (x, y) grid that scipy's
 *         LinearNDInterpolator builds (diag[(nx-1)*(ny-1)]: 0 = diagonal (0,0)-(1,1), 1 = (1,0)-(0,1)),
 *         0 outside the grid; y coordinate per depth in depth_y[d]; value * depth_scale[d]. */
typedef struct sd_table {
    int32_t kind;
    int32_t nx, ny;
    const double *x;           /* [nx] ascending, wavelength in Angstrom */
    const double *y;           /* [ny] ascending (kind 2) */
    const double *values;      /* [nx*ny] row-major (x major) */
    const uint8_t *diag;       /* kind 2 */
    const double *depth_y;     /* [D] (kind 2) */
    const double *depth_scale; /* [D] number density x unit scaling */
} sd_table;

#define SD_MAX_TABLES 8
typedef struct sd_continuum {
    /* hydrogenic bound-free (calc_alpha_bf :178-239): levels of all configured species, sorted by
     * cutoff; bf_prefix[(k)*D + d] = sum over the first k levels of BF*(ion+1)^4 n_level[d]/n^5 */
    int32_t n_bf_levels;
    const double *bf_nu_cut;   /* [n] ascending */
    const double *bf_prefix;   /* [(n+1)*D] */
    const double *ff_coef;     /* [D] sum_species FF Z^2 n_e n_ion / sqrt(T) (calc_alpha_ff :274-317) or NULL */
    const double *ray_c4, *ray_c6, *ray_c8; /* [D] Rayleigh coefficients x number densities (:111-125) or NULL */
    const double *electron;    /* [D] sigma_T n_e (calc_alpha_electron :139-174) or NULL */
    int32_t n_tables;
    sd_table tables[SD_MAX_TABLES]; /* calc_alpha_file :40-70, in opacity_config.file order */
} sd_continuum;

/* store_mask bit i: also keep source i as its own (D, p1-p0) array (for opacities_dict):
 * bit 0 bf, 1 ff, 2 rayleigh, 3 electron, 4.. tables.  total = sum of all sources + alpha_line slots
 * that were computed since the last sd_set_grid (Opacities.calc_total_alphas, opacities/base.py:24-28). */
#define SD_SRC_BF 0
#define SD_SRC_FF 1
#define SD_SRC_RAYLEIGH 2
#define SD_SRC_ELECTRON 3
#define SD_SRC_TABLE0 4
int sd_calc_continuum(sd_ctx *ctx, const sd_continuum *desc, uint32_t store_mask);

/* ---- K4: formal solution (raytrace radiation_field_solvers/base.py:271-346; single_theta_trace_parallel
 *          :85-268; calc_weights_parallel :6-47; blackbody_flux_at_nu source_functions/blackbody.py:11-35) */
/* ray_ds is (D-1, n_theta) row-major path lengths (plane-parallel dr/cos(theta) or calculate_spherical_ray
 * :349-381, both formed by the caller from O(D*n_theta) numbers); F_nu = scale * sum_theta w I_theta for
 * every depth; I_nus (D, W, n_theta) kept when track != 0.  Reads the TOTAL buffer. */
int sd_raytrace(sd_ctx *ctx, int32_t n_theta, const double *ray_ds, const double *weights, int32_t inward_rays,
                double scale, int32_t track);

/* ---- results ----------------------------------------------------------------------------------- */
#define SD_BUF_GAMMAS 1          /* (L, D) */
#define SD_BUF_DOPPLER 2         /* (L, D) */
#define SD_BUF_ALPHA_LINE 3      /* (D, W) slot 0 */
#define SD_BUF_ALPHA_MOLECULE 4  /* (D, W) slot 1 */
#define SD_BUF_TOTAL 5           /* (D, W) */
#define SD_BUF_F_NU 6            /* (D, W) */
#define SD_BUF_I_NUS 7           /* (D, W, n_theta) */
#define SD_BUF_LINE_STRENGTH 8   /* (L, D) alpha_line of the line table (uploaded or sd_calc_alpha_line_vald) */
#define SD_BUF_NUS 9             /* (1, N) the global frequency grid as uploaded by sd_set_grid */
#define SD_BUF_SOURCE0 16        /* + source index: (D, W) */
/* Copy a result into dst (host or device); count = number of doubles, must equal the buffer size. */
int sd_get(sd_ctx *ctx, int32_t which, double *dst, int64_t count);
/* Copy only row `row` of a (rows, W) buffer (e.g. the emergent spectrum F_nu[-1]). */
int sd_get_row(sd_ctx *ctx, int32_t which, int32_t row, double *dst, int64_t count);
/* Overwrite the TOTAL buffer (raytrace() on caller-supplied opacities, as the reference's fixtures do). */
int sd_set_total(sd_ctx *ctx, const double *total, int64_t count);
/* Device address of a result (zero-copy wrapping / NCCL all_gather); elements in *count. */
int sd_buffer(sd_ctx *ctx, int32_t which, void **device_ptr, int64_t *count);

/* ---- elementwise kernels (the reference's numba.cuda twins: voigt.py:94-195, broadening.py:74-547) -- */
int sd_ew_faddeeva(sd_ctx *ctx, int64_t n, const double *z_re, const double *z_im, double *w_re, double *w_im);
int sd_ew_voigt_profile(sd_ctx *ctx, int64_t n, const double *delta_nu, const double *doppler_width,
                        const double *gamma, double *phi);
int sd_ew_doppler_width(sd_ctx *ctx, int64_t n, const double *nu_line, const double *T, const double *mass,
                        double vmic, double *out);
int sd_ew_n_effective(sd_ctx *ctx, int64_t n, const double *z_eff, const double *e_ion, const double *e_level,
                      double *out);
int sd_ew_gamma_linear_stark(sd_ctx *ctx, int64_t n, const double *n_up, const double *n_lo, const double *n_e,
                             double *out);
int sd_ew_gamma_quadratic_stark(sd_ctx *ctx, int64_t n, const double *z_eff, const double *n_up,
                                const double *n_lo, const double *n_e, const double *T, double *out);
int sd_ew_gamma_van_der_waals(sd_ctx *ctx, int64_t n, const double *z_eff, const double *n_up, const double *n_lo,
                              const double *T, const double *n_H, double *out);
int sd_ew_blackbody(sd_ctx *ctx, int32_t n_depth, int64_t n_nu, const double *nus, const double *T, double *out);
int sd_ew_calc_weights(sd_ctx *ctx, int64_t n, const double *tau, double *w0, double *w1, double *w2);

/* ---- spectrum post-processing: rotation_broadening (broadening.py:824-877) ----------------------------------- */
/* out[i] = sum_j weights[j] x[reflect(i - j + m/2)], m odd: scipy.ndimage.convolve1d(x, weights) with its default
 * "reflect" boundary, the convolution rotation_broadening applies to the flux (the O(m) rotational profile itself is
 * formed by the caller).  x / out: n doubles, host or device. */
int sd_convolve1d_reflect(sd_ctx *ctx, int64_t n, const double *x, int32_t m, const double *weights, double *out);

/* ---- measurement helpers (bench.py roofline denominators) --------------------------------------- */
/* Dependent-chain-free DFMA loop on every SM; returns achieved FP64 TFLOP/s (2 flops per FMA). */
int sd_bench_dfma(sd_ctx *ctx, int32_t iters, double *tflops);
/* time of the device work enqueued between the two calls, in ms, measured with CUDA events on the
 * context's stream (sd_timer_stop synchronises). */
/* FP64 issue-rate probes (DFMA TFLOP/s): mode 0 = constant second/third operands (same as sd_bench_dfma),
 * 1 = three distinct register operands per DFMA, 2 = two distinct, 3 = mode 0 + one MUFU.RCP64H per 8 DFMA */
int sd_bench_fp64(sd_ctx *ctx, int32_t mode, int32_t iters, double *tflops);
/* the far-wing evaluation of k_lines in isolation (registers only): Gevals/s.  variant 0 = full body, 1 = no
 * reciprocal, 2 = MUFU seed without the Newton step */
int sd_bench_fareval(sd_ctx *ctx, int32_t variant, int32_t iters, double *gevals);
/* accuracy probe of the reciprocal used in the far-wing loop: MUFU.RCP64H seed, seed + one Newton step, seed + one
 * cubic step (tests only) */
int sd_debug_rcp(sd_ctx *ctx, int64_t n, const double *x, double *seed, double *quad, double *cubic);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
int64_t sd_launch_count(const sd_ctx *ctx);
int sd_timer_start(sd_ctx *ctx);
int sd_timer_stop(sd_ctx *ctx, float *ms);
/* Device time [ms] of the most recent run of each kernel group, from CUDA events the library records on the context's
 * stream around the launches (synchronises on them); -1 for a group that has not run.  out_ms[0] K1 k_broadening,
 * [1] K2 preparation (records, windows, class lists), [2] K2 edge sorts, [3] K2 k_far_coeffs (+ reduce, all levels),
 * [4] K2 k_lines, [5] K3 k_continuum, [6] K4 k_raytrace, [7] line-strength producer (sd_calc_alpha_line_*). */
#define SD_N_PHASE_TIMES 8
int sd_phase_times(sd_ctx *ctx, float out_ms[SD_N_PHASE_TIMES]);

#ifdef __cplusplus
}
#endif
#endif /* STARDIS_B200_H */
