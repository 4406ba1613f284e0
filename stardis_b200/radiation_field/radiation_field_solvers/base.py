"""Formal solution of radiative transfer on the device (stardis/radiation_field/radiation_field_solvers/base.py).

``raytrace`` (base.py:271-346) keeps the reference's signature and side effects: it reads
``stellar_radiation_field.opacities.total_alphas``, accumulates into ``F_nu`` and fills ``I_nus`` when intensities
are tracked.  All angles are solved in ONE kernel (csrc/k4_raytrace.cu) instead of the reference's serial loop over
angles."""
from __future__ import annotations

import numpy as np

from ... import _lib as L
from ... import units as u
from ...device import default_context
from ...device_array import DeviceArray, as_host


def calc_weights_parallel(delta_tau):
    """w0, w1, w2 of van Noort 2001 eq. 14 (base.py:6-47), elementwise on the device."""
    return default_context().calc_weights(np.asarray(delta_tau, dtype=np.float64))


calc_weights = calc_weights_parallel  # numpy twin of the reference (base.py:50-82): same values


def calculate_spherical_ray(thetas, depth_points_radii):
    """Path length of every ray through every shell in spherical geometry (base.py:349-381): O(D * n_theta) numbers,
    formed on the host and handed to the kernel.  Shells a ray does not reach give 0."""
    thetas = np.asarray(thetas, dtype=np.float64)
    r = np.asarray(u.values_of(depth_points_radii), dtype=np.float64)
    out = np.zeros((len(r) - 1, len(thetas)))
    for t, theta in enumerate(thetas):
        b = r[-1] * np.sin(theta)
        with np.errstate(invalid="ignore"):
            z = np.sqrt(r**2 - b**2)
        dz = np.diff(z)
        ok = ~np.isnan(dz)
        out[ok, t] = dz[ok]
    return out


def ray_distances(stellar_model, thetas):
    """(D-1, n_theta) path lengths and the inward-ray flag (base.py:296-306)."""
    if stellar_model.spherical:
        return calculate_spherical_ray(thetas, stellar_model.geometry.r), True
    dist = np.asarray(stellar_model.geometry.dist_to_next_depth_point, dtype=np.float64)
    return dist.reshape(-1, 1) / np.cos(np.asarray(thetas, dtype=np.float64)), False


def single_theta_trace_parallel(ray_dist_to_next_depth_point, temps, alphas, tracing_nus, source_function=None,
                                inward_rays=False):
    """One angle (base.py:85-268) -> I (D, N).  ``source_function`` is accepted for signature compatibility; the
    kernel evaluates the Planck function inline (the only source function the reference has)."""
    ctx = default_context()
    ctx.evict()
    alphas = np.ascontiguousarray(alphas, dtype=np.float64)
    ctx.set_atmosphere(np.ravel(u.values_of(temps)))
    ctx.set_grid(u.values_of(tracing_nus))
    ctx.set_total(alphas)
    ctx.raytrace(np.asarray(ray_dist_to_next_depth_point, dtype=np.float64).reshape(-1, 1), np.array([1.0]),
                 inward_rays=inward_rays)
    return ctx.get(L.BUF_F_NU)


def raytrace(stellar_model, stellar_radiation_field):
    """base.py:271-346: F_nu += sum_theta w_theta I_theta (every depth), x (r[-1]/reference_r)^2 when spherical."""
    srf = stellar_radiation_field
    ctx = getattr(srf, "device_context", None) or default_context()
    nus = u.values_of(srf.frequencies)
    N = nus.shape[0]
    D = stellar_model.no_of_depth_points
    shard = getattr(srf, "shard", None)
    p0, p1 = (0, N) if shard is None else (int(shard[0]), int(shard[1]))
    W = p1 - p0
    token = getattr(srf, "token", None)
    if token is None or ctx.owner != token:
        # the context does not hold this field's opacities (hand-filled total_alphas, or the context was reused)
        total = np.ascontiguousarray(np.asarray(srf.opacities.total_alphas), dtype=np.float64)
        ctx.evict()
        ctx.set_atmosphere(u.values_of(stellar_model.temperatures))
        ctx.set_grid(nus, p0, p1)
        ctx.set_total(total)
        ctx.owner = token
    ds, inward = ray_distances(stellar_model, srf.thetas)
    scale = 1.0
    if stellar_model.spherical:
        r = u.values_of(stellar_model.geometry.r)
        scale = float((r[-1] / float(u.cgs_values_of(stellar_model.geometry.reference_r))) ** 2)
    track = bool(getattr(srf, "track_individual_intensities", False))
    previous = srf._F_nu if hasattr(srf, "_F_nu") else getattr(srf, "F_nu", None)
    prev_host = None
    if previous is not None:
        arr = as_host(previous)  # materialises a device-backed F_nu BEFORE its buffer is overwritten
        if np.any(arr != 0):
            prev_host = arr
    held = getattr(srf, "_I_nus", None)
    if isinstance(held, DeviceArray):
        held.detach_to_host()  # intensities of an earlier raytrace() on this field: materialise before the buffer is reused
    ctx.raytrace(ds, np.asarray(srf.I_nus_weights, dtype=np.float64), inward_rays=inward, scale=scale, track=track)
    F = ctx.track(DeviceArray(ctx, L.BUF_F_NU, (D, W)))
    if prev_host is not None:
        F = (prev_host * scale) + F.numpy()  # `F_nu +=` then `*= correction` (base.py:336-344)
    srf.F_nu = F
    if track:
        srf.I_nus = ctx.track(DeviceArray(ctx, L.BUF_I_NUS, (D, W, len(srf.thetas))))
    return srf.F_nu
