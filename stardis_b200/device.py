"""Thin object wrapper over the C ABI: one ``DeviceContext`` per GPU / rank.

All arithmetic happens in libstardis_b200.so; this class only marshals pointers.  Inputs may be numpy arrays
(pageable host memory), pinned torch tensors or CUDA torch tensors -- the library resolves the direction of
every copy itself (UVA).  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib as L


class DeviceContext:
    def __init__(self, device: int = 0, stream=None):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.sd_create(C.byref(h), int(device))
        if rc != L.SD_OK:
            raise L.StardisB200Error(
                f"sd_create(device={device}) failed with code {rc}: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = int(device)
        self._keep = []  # host buffers that must outlive in-flight async copies
        self.D = 0
        self.N = 0
        self.p0 = self.p1 = 0
        self.L = 0
        self.n_theta = 0
        self._live = []  # weak references to DeviceArray results that still point into this context's buffers
        self.owner = None               # token of the RadiationField whose opacities/flux the buffers hold
        if stream is not None:
            self.set_stream(stream)

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != L.SD_OK:
            raise L.StardisB200Error(f"libstardis_b200 error {rc}: {self.lib.sd_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.sd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: raw cudaStream_t (int) or a torch.cuda.Stream; None restores the context's own stream."""
        raw = getattr(stream, "cuda_stream", stream)
        if stream is not None and not raw:
            raw = 1  # torch's default stream has handle 0 == cudaStreamLegacy; NULL would select the context's own stream
        self._ck(self.lib.sd_set_stream(self.h, raw))

    def synchronize(self):
        self._ck(self.lib.sd_synchronize(self.h))
        self._keep.clear()

    KEEP_LIMIT = 4096  # host buffers kept alive for in-flight copies before a synchronisation point is forced

    def _in(self, a, integer=False):
        if a is None:
            return None
        a = L.i64(a) if integer else L.f64(a)
        if len(self._keep) >= self.KEEP_LIMIT:  # callers that never synchronise (device-to-device result reads) must
            self.synchronize()                   # not grow the list without bound
        self._keep.append(a)
        return L.ptr(a)

    @property
    def W(self):
        return self.p1 - self.p0

    def track(self, device_array):
        self._live.append(weakref.ref(device_array))
        return device_array

    def evict(self):
        """Copy every still-referenced, not yet materialised result to the host: called before the buffers are
        reused for another radiation field."""
        for ref in self._live:
            a = ref()
            if a is not None:
                a.detach_to_host()
        self._live.clear()
        self.owner = None

    # ------------------------------------------------------------------ inputs
    def set_atmosphere(self, T, n_e=None, n_HI=None, vmic_cgs=0.0):
        T = L.f64(T)
        self.D = int(T.shape[0])
        self._ck(self.lib.sd_set_atmosphere(self.h, self.D, self._in(T), self._in(n_e), self._in(n_HI), float(vmic_cgs)))

    def set_grid(self, nus, p0=0, p1=None):
        nus = L.f64(nus)
        self.N = int(nus.shape[0])
        self.p0, self.p1 = int(p0), int(self.N if p1 is None else p1)
        self._ck(self.lib.sd_set_grid(self.h, self.N, self._in(nus), self.p0, self.p1))

    def set_grid_from(self, other, p0=0, p1=None):
        """The grid another context of this device already holds (device-to-device copy instead of a second upload)."""
        ptr_, n = other.buffer(L.BUF_NUS)
        self.N = int(n)
        self.p0, self.p1 = int(p0), int(self.N if p1 is None else p1)
        self._ck(self.lib.sd_set_grid(self.h, self.N, ptr_, self.p0, self.p1))

    def set_lines(self, nu, alpha_line, mass=None, atomic_number=None, ion_number=None, ionization_energy=None,
                  level_energy_upper=None, level_energy_lower=None, A_ul=None, stark=None, waals=None):
        s = L.SdLines()
        self.L = int(nu.shape[0])
        s.n_lines = self.L
        s.nu = self._in(nu)
        s.alpha_line = self._in(alpha_line)
        s.mass = self._in(mass)
        s.atomic_number = self._in(atomic_number, integer=True)
        s.ion_number = self._in(ion_number, integer=True)
        s.ionization_energy = self._in(ionization_energy)
        s.level_energy_upper = self._in(level_energy_upper)
        s.level_energy_lower = self._in(level_energy_lower)
        s.A_ul = self._in(A_ul)
        s.stark = self._in(stark)
        s.waals = self._in(waals)
        self._ck(self.lib.sd_set_lines(self.h, C.byref(s)))

    # ------------------------------------------------------------------ kernels
    def calc_broadening(self, flags):
        self._ck(self.lib.sd_calc_broadening(self.h, int(flags)))

    def set_broadening(self, gammas, doppler_widths):
        g = L.f64(gammas)
        cols = 1 if g.ndim == 1 else int(g.shape[1])
        self._ck(self.lib.sd_set_broadening(self.h, self._in(g), cols, self._in(doppler_widths)))

    def calc_alpha_line(self, slot=0):
        self._ck(self.lib.sd_calc_alpha_line(self.h, int(slot)))

    def calc_alpha_line_vald(self, n_over_u, ion_row, gf, e_low_erg, g_lo=None):
        """Line strengths (L, D) of a VALD linelist on the device (plasma/base.py:178-455); see include/stardis_b200.h.
        Needs ``set_atmosphere`` and ``set_lines(nu, None, ...)`` first; the result feeds K2 without a host round trip
        (``get(BUF_LINE_STRENGTH)`` copies it back)."""
        t = L.f64(n_over_u)
        if t.ndim != 2 or t.shape[1] != self.D:
            raise ValueError("n_over_u must have shape (n_ions, D)")
        rows = np.ascontiguousarray(ion_row, dtype=np.int64)
        if rows.shape[0] != self.L or (rows.size and (rows.min() < 0 or rows.max() >= t.shape[0])):
            raise ValueError("ion_row must hold one valid row of n_over_u per line")
        self._ck(self.lib.sd_calc_alpha_line_vald(self.h, int(t.shape[0]), self._in(t), self._in(rows, integer=True),
                                                  self._in(gf), self._in(g_lo), self._in(e_low_erg)))  # e_low None: E_lower column

    def calc_alpha_line_levels(self, level_number_density, g, lower_level_index, upper_level_index, f_lu,
                               metastable_upper=None):
        """Line strengths (L, D) of the tardis line list on the device (AlphaLine, plasma/base.py:130-175, with tardis'
        stimulated emission factor); see include/stardis_b200.h.  Needs ``set_atmosphere`` and ``set_lines(nu, None, ...)``."""
        t = L.f64(level_number_density)
        if t.ndim != 2 or t.shape[1] != self.D:
            raise ValueError("level_number_density must have shape (n_levels, D)")
        lo = np.ascontiguousarray(lower_level_index, dtype=np.int64)
        up = np.ascontiguousarray(upper_level_index, dtype=np.int64)
        for idx in (lo, up):
            if idx.shape[0] != self.L or (idx.size and (idx.min() < 0 or idx.max() >= t.shape[0])):
                raise ValueError("level indices must hold one valid row of level_number_density per line")
        meta = None if metastable_upper is None else np.ascontiguousarray(metastable_upper, dtype=np.int64)
        self._ck(self.lib.sd_calc_alpha_line_levels(self.h, int(t.shape[0]), self._in(t), self._in(g), self._in(lo, integer=True),
                                                    self._in(up, integer=True), self._in(meta, integer=True), self._in(f_lu)))

    def nonfinite_line_strengths(self):
        """Number of NaN/inf values the last device line-strength producer wrote (synchronises)."""
        out = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.sd_line_stats_ex(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return int(out[11])

    def set_farfield(self, on=True):
        """Far-field (Taylor) expansion of distant region-I wings per pixel tile; off = evaluate every pixel directly."""
        self._ck(self.lib.sd_set_farfield(self.h, int(bool(on))))

    def set_line_stats(self, on=True):
        self._ck(self.lib.sd_set_line_stats(self.h, int(bool(on))))

    def line_stats(self):
        out = np.zeros(8, dtype=np.int64)
        self._ck(self.lib.sd_line_stats(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(region_evals=out[:4].copy(), evals=int(out[:4].sum()), pairs=int(out[4]), wide_pairs=int(out[5]),
                    zero_doppler_pairs=int(out[6]))

    def line_stats_ex(self):
        """Executed work of the last counting pass (see sd_line_stats_ex in include/stardis_b200.h)."""
        out = np.zeros(16, dtype=np.int64)
        self._ck(self.lib.sd_line_stats_ex(self.h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(direct_region_evals=out[:4].copy(), far_replaced_evals=int(out[8]), far_expansions=int(out[9]),
                    far_terms=int(out[10]), multipole_expansions=int(out[12]), m2l_row_steps=int(out[13]))

    PHASES = ("K1_broadening", "K2_prepare", "K2_edge_sort", "K2_far_coeffs", "K2_lines", "K3_continuum", "K4_raytrace",
              "line_strengths")

    def phase_times(self):
        """Device time [ms] of the most recent run of every kernel group (CUDA events inside the library)."""
        out = (C.c_float * 8)()
        self._ck(self.lib.sd_phase_times(self.h, out))
        return {name: float(out[k]) for k, name in enumerate(self.PHASES)}

    def calc_continuum(self, bf_nu_cut=None, bf_prefix=None, ff_coef=None, rayleigh=None, electron=None, tables=(),
                       store_mask=0):
        """tables: sequence of dicts(kind, x, y, values, diag, depth_y, depth_scale)."""
        d = L.SdContinuum()
        d.n_bf_levels = 0 if bf_nu_cut is None else int(len(bf_nu_cut))
        d.bf_nu_cut = self._in(bf_nu_cut) if d.n_bf_levels else None
        d.bf_prefix = self._in(bf_prefix)
        d.ff_coef = self._in(ff_coef)
        if rayleigh is not None:
            d.ray_c4, d.ray_c6, d.ray_c8 = (self._in(x) for x in rayleigh)
        d.electron = self._in(electron)
        if len(tables) > L.MAX_TABLES:
            raise ValueError(f"at most {L.MAX_TABLES} file opacity tables are supported")
        d.n_tables = len(tables)
        for k, t in enumerate(tables):
            tb = d.tables[k]
            tb.kind = int(t["kind"])
            tb.nx = int(len(t["x"]))
            tb.ny = int(len(t["y"])) if t.get("y") is not None else 1
            tb.x = self._in(t["x"])
            tb.y = self._in(t.get("y"))
            tb.values = self._in(t["values"])
            if t.get("diag") is not None:
                dg = np.ascontiguousarray(t["diag"], dtype=np.uint8)
                self._keep.append(dg)
                tb.diag = dg.ctypes.data
            tb.depth_y = self._in(t.get("depth_y"))
            tb.depth_scale = self._in(t["depth_scale"])
        self._ck(self.lib.sd_calc_continuum(self.h, C.byref(d), int(store_mask)))

    def set_total(self, total):
        t = L.f64(total)
        self._ck(self.lib.sd_set_total(self.h, self._in(t), int(np.prod(t.shape))))

    def raytrace(self, ray_ds, weights, inward_rays=False, scale=1.0, track=False):
        ds = L.f64(ray_ds)
        w = L.f64(weights)
        self.n_theta = int(w.shape[0])
        self._ck(self.lib.sd_raytrace(self.h, self.n_theta, self._in(ds), self._in(w), int(bool(inward_rays)), float(scale),
                                      int(bool(track))))

    # ------------------------------------------------------------------ results
    def _shape(self, which):
        if which in (L.BUF_GAMMAS, L.BUF_DOPPLER, L.BUF_LINE_STRENGTH):
            return None
        if which == L.BUF_I_NUS:
            return (self.D, self.W, self.n_theta)
        return (self.D, self.W)

    def get(self, which, out=None, shape=None):
        """Copy a result buffer into ``out`` (numpy array or torch tensor, host or device); synchronises when
        ``out`` is a host array that this call allocated."""
        if out is None:
            shape = shape or self._shape(which)
            if shape is None:
                ptr_, cnt = self.buffer(which)
                shape = (self.L, cnt // max(self.L, 1))
            out = np.empty(shape, dtype=np.float64)
            self._ck(self.lib.sd_get(self.h, int(which), L.ptr(out), int(out.size)))
            self.synchronize()
            return out
        n = out.numel() if hasattr(out, "numel") else out.size
        self._ck(self.lib.sd_get(self.h, int(which), L.ptr(out), int(n)))
        return out

    def get_row(self, which, row, out=None):
        if out is None:
            out = np.empty(self.W, dtype=np.float64)
            self._ck(self.lib.sd_get_row(self.h, int(which), int(row), L.ptr(out), int(out.size)))
            self.synchronize()
            return out
        n = out.numel() if hasattr(out, "numel") else out.size
        self._ck(self.lib.sd_get_row(self.h, int(which), int(row), L.ptr(out), int(n)))
        return out

    def buffer(self, which):
        p = C.c_void_p()
        n = C.c_int64()
        self._ck(self.lib.sd_buffer(self.h, int(which), C.byref(p), C.byref(n)))
        return p.value, n.value

    # ------------------------------------------------------------------ elementwise twins
    def _ew(self, fn, ins, n_out, extra=()):
        ins = [np.ascontiguousarray(x, dtype=np.float64) for x in np.broadcast_arrays(*[np.asarray(x, dtype=np.float64) for x in ins])]
        shape = ins[0].shape
        n = int(ins[0].size)
        outs = [np.empty(shape, dtype=np.float64) for _ in range(n_out)]
        args = [self.h, n] + [L.ptr(x) for x in ins] + list(extra) + [L.ptr(o) for o in outs]
        self._ck(fn(*args))
        self.synchronize()
        return outs

    def faddeeva(self, z):
        z = np.asarray(z, dtype=np.complex128)
        wr, wi = self._ew(self.lib.sd_ew_faddeeva, [z.real, z.imag], 2)
        return wr + 1j * wi

    def voigt_profile(self, delta_nu, doppler_width, gamma):
        return self._ew(self.lib.sd_ew_voigt_profile, [delta_nu, doppler_width, gamma], 1)[0]

    def doppler_width(self, nu_line, T, mass, vmic):
        return self._ew(self.lib.sd_ew_doppler_width, [nu_line, T, mass], 1, extra=(C.c_double(float(vmic)),))[0]

    def n_effective(self, z_eff, e_ion, e_level):
        return self._ew(self.lib.sd_ew_n_effective, [z_eff, e_ion, e_level], 1)[0]

    def gamma_linear_stark(self, n_up, n_lo, n_e):
        return self._ew(self.lib.sd_ew_gamma_linear_stark, [n_up, n_lo, n_e], 1)[0]

    def gamma_quadratic_stark(self, z_eff, n_up, n_lo, n_e, T):
        return self._ew(self.lib.sd_ew_gamma_quadratic_stark, [z_eff, n_up, n_lo, n_e, T], 1)[0]

    def gamma_van_der_waals(self, z_eff, n_up, n_lo, T, n_H):
        return self._ew(self.lib.sd_ew_gamma_van_der_waals, [z_eff, n_up, n_lo, T, n_H], 1)[0]

    def calc_weights(self, tau):
        return tuple(self._ew(self.lib.sd_ew_calc_weights, [tau], 3))

    def blackbody(self, nus, T):
        nus = np.ascontiguousarray(nus, dtype=np.float64)
        T = np.ascontiguousarray(np.ravel(T), dtype=np.float64)
        out = np.empty((T.size, nus.size))
        self._ck(self.lib.sd_ew_blackbody(self.h, int(T.size), int(nus.size), L.ptr(nus), L.ptr(T), L.ptr(out)))
        self.synchronize()
        return out

    def convolve1d_reflect(self, x, weights):
        """scipy.ndimage.convolve1d(x, weights) (mode "reflect", odd kernel) on the device."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        out = np.empty_like(x)
        self._ck(self.lib.sd_convolve1d_reflect(self.h, int(x.size), L.ptr(x), int(w.size), L.ptr(w), L.ptr(out)))
        self.synchronize()
        return out

    # ------------------------------------------------------------------ measurement
    def bench_dfma(self, iters=4096):
        t = C.c_double()
        self._ck(self.lib.sd_bench_dfma(self.h, int(iters), C.byref(t)))
        return t.value

    def launch_count(self):
        return int(self.lib.sd_launch_count(self.h))

    def timer_start(self):
        self._ck(self.lib.sd_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.lib.sd_timer_stop(self.h, C.byref(ms)))
        return ms.value


_default = {}


def default_context(device: int = 0) -> DeviceContext:
    """Process-wide context of a device (created on first use; raises without a GPU)."""
    if device not in _default:
        _default[device] = DeviceContext(device)
    return _default[device]
