"""Config + model parsing for ``run_stardis`` (behaviour of stardis/io/base.py:20-132, hot-path scope).

Atomic data and the LTE plasma live in the third-party ``tardis`` package (not available offline, out of the hot-path
scope).  ``atom_data`` therefore accepts

* ``*.h5``  -- a carsus/tardis atomic data file: NOT handled here.  Turning it into the hot path's inputs is the job of
  tardis' LTE plasma and of the reference's ``create_stellar_plasma`` (upstream of the path, third-party, absent
  offline).  A user with tardis + stardis installed builds ``stellar_model`` / ``stellar_plasma`` with the reference's
  own ``parse_config_to_model`` / ``create_stellar_plasma`` and hands them to
  ``stardis_b200.radiation_field.create_stellar_radiation_field``; ``run_stardis`` raises
  ``NotImplementedError`` for such a config instead of pretending to be a drop-in for that part;
* ``synthetic:<n_lines>[:<seed>]`` -- the seeded synthetic plasma state of ``stardis_b200.plasma.synthetic`` (the
  attribute surface the hot path reads), used by the benchmarks and tests.
"""
from __future__ import annotations

import logging
from pathlib import Path

from .. import units as u
from .config import Configuration, load_config, validate_config
from .model.marcs import read_marcs_model

logger = logging.getLogger(__name__)


def parse_config_to_model(config_fname, add_config_dict=None):
    """-> (config, adata, stellar_model).  ``adata`` is a tardis AtomData for ``.h5`` inputs, else a small dict
    describing the synthetic plasma request."""
    try:
        config = load_config(config_fname)
    except Exception as e:
        raise ValueError("Config failed to validate. Check the config file.") from e
    if add_config_dict:
        logger.info("Updating config with additional keys and values")
        for key, val in add_config_dict.items():
            try:
                config.set_config_item(key, val)
            except Exception as e:
                raise ValueError(f"{key} not a valid type. Should be a string for keys.") from e
        try:
            config = Configuration(validate_config(config))
        except Exception as e:
            raise ValueError("Additional config keys and values failed to validate.") from e

    base = Path(config_fname).resolve().parent
    adata = _load_atom_data(config.atom_data, base)
    if config.input_model.get("nuclide_rescaling_dict"):
        raise NotImplementedError(
            "input_model.nuclide_rescaling_dict rescales the tardis composition before the plasma solve (stardis/io/base.py:"
            "103-117), which is upstream of the hot path and not part of this package; it would be silently ignored here")

    logger.info("Reading model")
    if config.input_model.type == "marcs":
        fname = Path(config.input_model.fname)
        if not fname.is_absolute() and not fname.exists():
            fname = base / fname
        raw = read_marcs_model(fname, gzipped=config.input_model.gzipped)
        stellar_model = raw.to_stellar_model(
            adata if not isinstance(adata, dict) else None,
            final_atomic_number=config.input_model.final_atomic_number,
            composition_source=config.input_model.composition_source,
            helium_mass_frac_Y=config.input_model.composition_Y,
            heavy_metal_mass_frac_Z=config.input_model.composition_Z)
        stellar_model.raw_model = raw
        if config.opacity.line.disable_microturbulence:
            stellar_model.microturbulence = u.Quantity(0.0, u.km_s)
    elif config.input_model.type == "mesa":
        raise NotImplementedError("MESA model input belongs to the reference's IO layer (out of the hot-path scope)")
    else:
        raise ValueError("Model type not recognized. Must be either 'marcs' or 'mesa'")
    return config, adata, stellar_model


def _load_atom_data(spec, base):
    if spec.startswith("synthetic:"):
        parts = spec.split(":")
        return {"synthetic": True, "n_lines": int(parts[1]), "seed": int(parts[2]) if len(parts) > 2 else 0}
    raise NotImplementedError(
        f"atom_data={spec!r}: carsus/tardis atomic data files feed tardis' LTE plasma (stardis/plasma/base.py:491-569), "
        "which is upstream of the hot path and not part of this package.  With tardis + stardis installed, build "
        "stellar_model and stellar_plasma with the reference's own parse_config_to_model / create_stellar_plasma and pass "
        "them to stardis_b200.radiation_field.create_stellar_radiation_field (INTEGRATION.md); for benchmarks and tests use "
        "atom_data: 'synthetic:<n_lines>[:<seed>]'.")
