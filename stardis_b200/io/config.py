"""YAML configuration with the semantics of the reference's schemas (stardis/io/schemas/*.yml).

The reference validates with tardis' ``validate_yaml`` (JSON-schema draft-04 with default injection) and wraps the
result in tardis' ``Configuration`` (attribute-access dict).  tardis is not a dependency of the hot path, so the same
defaults, enums and ``additionalProperties`` rules are restated here in plain Python; when tardis is importable its
validator can be used instead by the reference's own IO layer.
"""
from __future__ import annotations

import copy

import yaml


class Configuration(dict):
    """dict with attribute access and dotted ``set_config_item`` (tardis.io.configuration.config_reader.Configuration)."""

    def __init__(self, value=None):
        super().__init__()
        for k, v in (value or {}).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, Configuration):
            value = Configuration(value)
        super().__setitem__(key, value)

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:
            raise AttributeError(item) from e

    def __setattr__(self, key, value):
        self[key] = value

    def set_config_item(self, config_item_string, value):
        if not isinstance(config_item_string, str):
            raise ValueError("config keys must be strings")
        keys = config_item_string.split(".")
        node = self
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = value

    def get_config_item(self, config_item_string):
        node = self
        for k in config_item_string.split("."):
            node = node[k]
        return node


_BROADENING = {"linear_stark", "quadratic_stark", "van_der_waals", "radiation"}
_RAYLEIGH = {"H", "He", "H2"}


def _only(d, allowed, where):
    extra = set(d) - set(allowed)
    if extra:
        raise ValueError(f"unknown key(s) {sorted(extra)} in {where}")


def _species_dict(d, where):
    out = {}
    for spec, opts in (d or {}).items():
        opts = dict(opts or {})
        _only(opts, ("gaunt", "departure"), f"{where}.{spec}")
        out[spec] = {"gaunt": opts.get("gaunt"), "departure": opts.get("departure")}
    return out


def validate_config(raw):
    """Defaults and checks of config_schema.yml, input_model.yml, opacity.yml, line.yml, result_options.yml."""
    raw = copy.deepcopy(dict(raw))
    if raw.get("stardis_config_version") != 1.0:
        raise ValueError("stardis_config_version must be 1.0")
    if not isinstance(raw.get("atom_data"), str):
        raise ValueError("atom_data (path) is required")
    cfg = {"stardis_config_version": 1.0, "atom_data": raw["atom_data"]}
    n_threads = raw.get("n_threads", 1)
    if int(n_threads) != n_threads:
        raise ValueError("n_threads must be an integer")
    cfg["n_threads"] = int(n_threads)
    no_of_thetas = raw.get("no_of_thetas", 10)
    if int(no_of_thetas) != no_of_thetas:
        raise ValueError("no_of_thetas must be an integer")
    cfg["no_of_thetas"] = int(no_of_thetas)

    # the reference's own test configs use the key "model" instead of "input_model" (SURVEY.md section 4): accept both
    im = dict(raw.get("input_model") or raw.get("model") or {})
    if im.get("type") not in ("marcs", "mesa"):
        raise ValueError("input_model.type must be 'marcs' or 'mesa'")
    if not isinstance(im.get("fname"), str):
        raise ValueError("input_model.fname is required")
    cfg["input_model"] = {
        "type": im["type"], "fname": im["fname"], "gzipped": bool(im.get("gzipped", False)),
        "final_atomic_number": int(im.get("final_atomic_number", 92)),
        "truncate_to_shell": int(im.get("truncate_to_shell", -99)),
        "composition_source": im.get("composition_source", "from_model"),
        "composition_Y": float(im.get("composition_Y", -99.0)), "composition_Z": float(im.get("composition_Z", -99.0)),
        "nuclide_rescaling_dict": dict(im.get("nuclide_rescaling_dict", {}) or {}),
    }

    op = dict(raw.get("opacity") or {})
    _only(op, ("file", "bf", "ff", "rayleigh", "disable_electron_scattering", "line"), "opacity")
    files = dict(op.get("file") or {})
    for k, v in files.items():
        if not (k.endswith("_bf") or k.endswith("_ff")) or not isinstance(v, str):
            raise ValueError(f"opacity.file key {k!r} must end in _bf or _ff and map to a path")
    rayleigh = list(op.get("rayleigh") or [])
    if set(rayleigh) - _RAYLEIGH:
        raise ValueError(f"opacity.rayleigh entries must be among {sorted(_RAYLEIGH)}")
    line = dict(op.get("line") or {})
    _only(line, ("disable", "broadening", "disable_microturbulence", "vald_linelist", "include_molecules"), "opacity.line")
    broadening = list(line.get("broadening") or [])
    if set(broadening) - _BROADENING:
        raise ValueError(f"opacity.line.broadening entries must be among {sorted(_BROADENING)}")
    vald = dict(line.get("vald_linelist") or {})
    cfg["opacity"] = {
        "file": files, "bf": _species_dict(op.get("bf"), "opacity.bf"), "ff": _species_dict(op.get("ff"), "opacity.ff"),
        "rayleigh": rayleigh, "disable_electron_scattering": bool(op.get("disable_electron_scattering", False)),
        "line": {
            "disable": bool(line.get("disable", False)), "broadening": broadening,
            "disable_microturbulence": bool(line.get("disable_microturbulence", False)),
            "vald_linelist": {"use_linelist": bool(vald.get("use_linelist", False)), "shortlist": bool(vald.get("shortlist", False)),
                              "use_vald_broadening": bool(vald.get("use_vald_broadening", True))},
            "include_molecules": bool(line.get("include_molecules", False)),
        },
    }
    ro = dict(raw.get("result_options") or {})
    _only(ro, ("return_model", "return_plasma", "return_radiation_field"), "result_options")
    cfg["result_options"] = {k: bool(ro.get(k, False)) for k in ("return_model", "return_plasma", "return_radiation_field")}
    return cfg


def load_config(config_fname):
    with open(config_fname) as fh:
        raw = yaml.safe_load(fh)
    return Configuration(validate_config(raw))
