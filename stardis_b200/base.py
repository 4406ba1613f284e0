"""Top-level API (stardis/base.py:13-141): ``run_stardis``, ``set_num_threads``, ``STARDISOutput``."""
from __future__ import annotations

import logging

import numpy as np

from . import units as u
from .io.base import parse_config_to_model
from .radiation_field.base import create_stellar_radiation_field

logger = logging.getLogger(__name__)

__all__ = ["run_stardis", "set_num_threads", "STARDISOutput"]


def run_stardis(config_fname, tracing_lambdas_or_nus, add_config_dict=None, device_context=None, shard=None):
    """Runs a STARDIS simulation (stardis/base.py:13-45).

    ``config_fname``: YAML configuration; ``tracing_lambdas_or_nus``: wavelengths or frequencies with units
    (``stardis_b200.units`` or astropy); ``add_config_dict``: dotted-key overrides.  Returns ``STARDISOutput``.
    Opacities and the formal solution run on the GPU; ``device_context`` / ``shard`` select the device and, for
    multi-GPU runs, the pixel range of the global grid this rank evaluates (``shard="auto"``: this rank's range of
    the cost-balanced partition, see ``distributed.line_balanced_bounds``)."""
    tracing_nus = u.to_hz(tracing_lambdas_or_nus)
    config, adata, stellar_model = parse_config_to_model(config_fname, add_config_dict)
    set_num_threads(config.n_threads)
    stellar_plasma = create_stellar_plasma(stellar_model, adata, config, tracing_nus)
    stellar_radiation_field = create_stellar_radiation_field(tracing_nus, stellar_model, stellar_plasma, config,
                                                             device_context=device_context, shard=shard)
    return STARDISOutput(config.result_options, stellar_model, stellar_plasma, stellar_radiation_field)


def create_stellar_plasma(stellar_model, adata, config, tracing_nus):
    """The LTE plasma is tardis' job (stardis/plasma/base.py:491-569) and outside this package: ``run_stardis`` covers the
    seeded synthetic provider (``atom_data: synthetic:<n_lines>``); real atomic data is refused in
    ``io.base.parse_config_to_model`` with a pointer to the integration route."""
    if isinstance(adata, dict) and adata.get("synthetic"):
        from .plasma.synthetic import create_synthetic_plasma

        raw = stellar_model.raw_model.data
        atm = dict(T=u.values_of(stellar_model.temperatures), pe=raw["pe"].values[::-1].astype(float),
                   pg=raw["pg"].values[::-1].astype(float))
        nus = u.values_of(tracing_nus)
        return create_synthetic_plasma(atm, adata["n_lines"], nus.min(), nus.max(), seed=adata["seed"],
                                       vald=config.opacity.line.vald_linelist.use_linelist)
    raise NotImplementedError("only the synthetic plasma provider is available (see io.base.parse_config_to_model)")


def set_num_threads(n_threads):
    """stardis/base.py:48-81.  The GPU path has no host thread pool; the accepted values and the error are the
    reference's (1, > 1 or -99; anything else -- including the 0 the schema documents -- raises ValueError)."""
    if n_threads == 1:
        logger.info("Running in serial mode")
    elif n_threads == -99:
        logger.info("Running with max threads")
    elif n_threads > 1:
        logger.info(f"Running with {n_threads} threads")
    else:
        raise ValueError(
            "n_threads must be a positive integer less than the number of available threads, or -99 to run with max threads.")


class STARDISOutput:
    """Key outputs of a run (stardis/base.py:84-141): optional ``stellar_model`` / ``stellar_plasma`` /
    ``stellar_radiation_field`` per ``result_options``; ``nus`` [Hz], ``lambdas`` [Angstrom]; ``spectrum_nu`` =
    ``F_nu[-1]`` [erg/s/cm^2/Hz]; ``spectrum_lambda`` = ``(F_nu nu / lambda)[-1]`` [erg/s/cm^2/Angstrom].

    Only the emergent row of ``F_nu`` is copied from the device.  In a multi-GPU run (``shard`` set and
    torch.distributed initialised) the spectrum shards are all-gathered so that every rank holds the full spectrum."""

    def __init__(self, result_options, stellar_model, stellar_plasma, stellar_radiation_field):
        if result_options.return_model:
            self.stellar_model = stellar_model
        if result_options.return_plasma:
            self.stellar_plasma = stellar_plasma
        if result_options.return_radiation_field:
            self.stellar_radiation_field = stellar_radiation_field

        self.nus = stellar_radiation_field.frequencies
        self.lambdas = u.Quantity(u.values_of(self.nus), u.Hz).to(u.AA, u.spectral())
        emergent = np.asarray(stellar_radiation_field.F_nu[-1], dtype=np.float64)
        shard = getattr(stellar_radiation_field, "shard", None)
        nus = u.values_of(self.nus)
        lambdas = self.lambdas.value
        if shard is not None and emergent.shape[0] != len(nus):
            from .distributed import allgather_spectrum, dist_info

            if dist_info()[0] is not None:
                emergent = allgather_spectrum(emergent, shard, len(nus),
                                              bounds=getattr(stellar_radiation_field, "shard_bounds", None))
            else:  # single process evaluating one shard: the spectrum covers pixels [p0, p1) only
                self.shard = (int(shard[0]), int(shard[1]))
                nus, lambdas = nus[shard[0]:shard[1]], lambdas[shard[0]:shard[1]]
        self.spectrum_nu = u.Quantity(emergent, "erg/s/cm2/Hz")
        # F_lambda [erg/s/cm^2/A] = F_nu * nu / lambda
        self.spectrum_lambda = u.Quantity(emergent * nus / lambdas, "erg/s/cm2/AA")
