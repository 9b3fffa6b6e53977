"""Config base class and versioned JSON (de)serialisation.

Mirrors the contract of torchao/core/config.py (AOBaseConfig :27-67, config_to_dict :174,
config_from_dict :204): a config is a dataclass with a ``version`` field; the wire format is
``{"_type": ClassName, "_version": v, "_data": {field: encoded}}`` so checkpoints carrying a
torchao quantization config can be read back.  Enums, torch.dtype, dataclasses (granularities,
Float8MMConfig) and nested configs are encoded recursively.
"""
from __future__ import annotations

import abc
import dataclasses
import enum
import importlib
import json
from typing import Any, Dict

import torch

__all__ = ["AOBaseConfig", "config_to_dict", "config_from_dict", "ConfigJSONEncoder", "ALLOWED_AO_MODULES"]


class AOBaseConfig(abc.ABC):
    """Base class of every workflow config; subclasses are dataclasses with a ``version``."""

    version: int = 1


# Modules searched, in order, for the class named by "_type" (the wire format stores bare class names, like the
# reference: torchao/core/config.py:174-305 looks names up in its own ALLOWED_AO_MODULES).  Nothing else is importable.
ALLOWED_AO_MODULES = (
    "ao_b200.quantization",
    "ao_b200.quantization.granularity",
    "ao_b200.quantization.quant_primitives",
    "ao_b200.quantization.quantize_.common",
    "ao_b200.quantization.quantize_.workflows",
    "ao_b200.float8.inference",
    "ao_b200.prototype.mx_formats",
)


class ConfigJSONEncoder(json.JSONEncoder):
    """Same wire format as the reference's encoder (torchao/core/config.py:70-172): configs, NamedTuples and dataclasses
    become ``{"_type": name, "_version": v, "_data": {...}}`` (configs: every public instance attribute except
    ``version``), enums ``{"_type": name, "_data": member name}``, dtypes ``{"_type": "torch.dtype", "_data": name}``;
    plain strings / numbers / None pass through (so a str-Enum field given as a string stays a string)."""

    def default(self, o):
        if isinstance(o, AOBaseConfig):
            data = {k: self.encode_value(v) for k, v in o.__dict__.items() if not k.startswith("_") and k != "version"}
            return {"_type": type(o).__name__, "_version": getattr(o, "version", 1), "_data": data}
        if isinstance(o, tuple) and hasattr(o, "_fields") and hasattr(o, "_asdict"):  # NamedTuple (Float8MMConfig)
            return {"_type": type(o).__name__, "_version": getattr(o, "version", 1),
                    "_data": {k: self.encode_value(v) for k, v in o._asdict().items()}}
        if dataclasses.is_dataclass(o) and not isinstance(o, type):
            return {"_type": type(o).__name__, "_version": getattr(o, "version", 1),
                    "_data": {f.name: self.encode_value(getattr(o, f.name)) for f in dataclasses.fields(o) if f.name != "version"}}
        if isinstance(o, enum.Enum):
            return {"_type": type(o).__name__, "_data": o.name}
        if isinstance(o, torch.dtype):
            return {"_type": "torch.dtype", "_data": str(o).split(".")[-1]}
        if isinstance(o, list):
            return [self.encode_value(v) for v in o]
        if isinstance(o, tuple):
            raise NotImplementedError(f"Tuples will be serialized as List in JSON, use Lists instead to avoid surprises. got: {o}")
        if isinstance(o, dict):
            return {k: self.encode_value(v) for k, v in o.items()}
        return super().default(o)

    def encode_value(self, value):
        try:
            return self.default(value)
        except TypeError:
            return value


def config_to_dict(config: AOBaseConfig) -> Dict[str, Any]:
    if not isinstance(config, AOBaseConfig):
        raise TypeError(f"expected an AOBaseConfig, got {type(config)}")
    return json.loads(json.dumps(config, cls=ConfigJSONEncoder))


def _resolve(name: str):
    for modname in ALLOWED_AO_MODULES:
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        obj = getattr(mod, name, None)
        if isinstance(obj, type):
            return obj
    raise ValueError(f"Failed to find class {name} in any of the allowed modules: {', '.join(ALLOWED_AO_MODULES)}")


def _decode(o: Any) -> Any:
    if isinstance(o, list):
        return [_decode(v) for v in o]
    if isinstance(o, dict):
        if "_type" in o and "_data" in o:
            t = o["_type"]
            if t == "torch.dtype":
                return getattr(torch, o["_data"])
            cls = _resolve(t)
            if issubclass(cls, enum.Enum):
                return cls[o["_data"]]
            kwargs = {k: _decode(v) for k, v in o["_data"].items()}
            version = o.get("_version", None)
            cur = getattr(cls, "version", 1)
            if version is not None and isinstance(cur, int) and version > cur:
                raise ValueError(f"{t}: stored version {version} is newer than supported version {cur}")
            fields = {f.name for f in dataclasses.fields(cls)} if dataclasses.is_dataclass(cls) else set()
            if "version" in fields and version is not None:
                kwargs["version"] = version
            return cls(**kwargs)
        return {k: _decode(v) for k, v in o.items()}
    return o


def config_from_dict(data: Dict[str, Any]) -> AOBaseConfig:
    if not isinstance(data, dict) or "_type" not in data or "_data" not in data:
        raise ValueError("config dict must carry '_type' and '_data'")
    out = _decode(data)
    if not isinstance(out, AOBaseConfig):
        raise ValueError(f"decoded object {type(out)} is not an AOBaseConfig")
    return out
