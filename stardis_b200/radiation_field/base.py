"""RadiationField container and driver (stardis/radiation_field/base.py:12-117)."""
from __future__ import annotations

import functools
import itertools
import logging

import numpy as np

from .. import units as u
from .opacities import Opacities
from .opacities.opacities_solvers import calc_alphas
from .radiation_field_solvers import raytrace
from .source_functions.blackbody import blackbody_flux_at_nu

logger = logging.getLogger(__name__)
_tokens = itertools.count(1)


@functools.lru_cache(maxsize=32)
def _leggauss(n):
    x, w = np.polynomial.legendre.leggauss(n)  # an eigenvalue problem: 0.2 ms, cached per angle count
    x.setflags(write=False)
    w.setflags(write=False)
    return x, w


class RadiationField:
    """Frequencies, opacities, fluxes and angle quadrature of one run (radiation_field/base.py:12-68).

    Attributes follow the reference: ``frequencies``, ``source_function``, ``opacities``, ``F_nu`` (D, N),
    ``thetas``, ``I_nus_weights``, ``track_individual_intensities`` and ``I_nus`` (D, N, n_theta) when tracked.
    ``F_nu`` / ``I_nus`` are allocated lazily (zeros) and replaced by device-backed arrays by ``raytrace``.

    B200 additions: ``device_context`` (None = the process-wide context of cuda:0) and ``shard`` = (p0, p1), the
    pixel range of the global grid this rank evaluates (None = everything); ``shard_bounds`` = the partition of all
    ranks when it is not the equal-width one (``distributed.line_balanced_bounds``); ``depth_shard`` = (rank, world):
    the opacity stages of this rank cover the depth points rank, rank + world, ... of the whole grid and are
    redistributed to the pixel ranges by one all-to-all before the formal solution (``stardis_b200.distributed``)."""

    hdf_properties = ["frequencies", "opacities", "F_nu"]

    def __init__(self, frequencies, source_function, stellar_model, num_of_thetas, track_individual_intensities=False,
                 device_context=None, shard=None, shard_bounds=None, depth_shard=None):
        self.frequencies = frequencies
        self.source_function = source_function
        self.opacities = Opacities(frequencies, stellar_model)
        self._shape = (stellar_model.no_of_depth_points, len(frequencies))
        self._F_nu = None
        # Gauss-Legendre nodes mapped as in the reference (radiation_field/base.py:60-63) -- deliberately NOT the
        # usual affine map onto [0, pi/2]
        thetas, weights = _leggauss(int(num_of_thetas))
        self.thetas = (thetas / 2) + 0.5 * np.pi / 2
        self.I_nus_weights = weights * np.pi / 2
        self.track_individual_intensities = track_individual_intensities
        self._I_nus = None
        self.device_context = device_context
        self.shard = shard
        self.shard_bounds = shard_bounds
        self.depth_shard = depth_shard
        self.token = next(_tokens)

    @property
    def F_nu(self):
        if self._F_nu is None:
            self._F_nu = np.zeros(self._shape)
        return self._F_nu

    @F_nu.setter
    def F_nu(self, value):
        self._F_nu = value

    @property
    def I_nus(self):
        if not self.track_individual_intensities:
            raise AttributeError("I_nus is only tracked when track_individual_intensities=True")
        if self._I_nus is None:
            self._I_nus = np.zeros(self._shape + (len(self.thetas),))
        return self._I_nus

    @I_nus.setter
    def I_nus(self, value):
        self._I_nus = value


def create_stellar_radiation_field(tracing_nus, stellar_model, stellar_plasma, config, device_context=None, shard=None):
    """radiation_field/base.py:71-117: RadiationField -> calc_alphas -> raytrace.

    Inside an initialised torch.distributed job, ``shard="auto"`` gives every rank the depth points rank, rank + world,
    ... for the opacity stages and an equal-width pixel range for the formal solution (one all-to-all in between, see
    ``stardis_b200.distributed``); ``shard="auto-nu"`` shards everything by frequency instead (cost-balanced ranges,
    ``distributed.line_balanced_bounds``; no exchange, but the per-(line, depth) preparation is repeated on every rank)."""
    shard_bounds, depth_shard = None, None
    if isinstance(shard, str):
        if shard not in ("auto", "auto-nu"):
            raise ValueError("shard must be None, (p0, p1), 'auto' or 'auto-nu'")
        from ..distributed import all_shards, dist_info, line_balanced_bounds

        mode = shard
        _, rank, world = dist_info()
        shard = None
        if world > 1 and mode == "auto":
            shard_bounds = all_shards(len(tracing_nus), world)
            shard = shard_bounds[rank]
            depth_shard = (rank, world)
        elif world > 1:
            from .opacities.opacities_solvers.base import select_lines

            line_nus = ([] if config.opacity.line.disable
                        else select_lines(stellar_plasma, stellar_model, tracing_nus, config.opacity.line).nu)
            shard_bounds = line_balanced_bounds(u.values_of(tracing_nus), line_nus, world)
            shard = shard_bounds[rank]
    stellar_radiation_field = RadiationField(
        tracing_nus, blackbody_flux_at_nu, stellar_model, config.no_of_thetas,
        track_individual_intensities=config.result_options.return_radiation_field,
        device_context=device_context, shard=shard, shard_bounds=shard_bounds, depth_shard=depth_shard)
    logger.info("Calculating alphas")
    calc_alphas(stellar_plasma=stellar_plasma, stellar_model=stellar_model,
                stellar_radiation_field=stellar_radiation_field, opacity_config=config.opacity,
                store_components=bool(config.result_options.return_radiation_field))
    logger.info("Raytracing")
    raytrace(stellar_model, stellar_radiation_field)
    return stellar_radiation_field
