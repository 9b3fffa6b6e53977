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


ALLOWED_AO_MODULES = {
    "ao_b200",
    "ao_b200.quantization",
    "ao_b200.quantization.granularity",
    "ao_b200.quantization.quant_primitives",
    "ao_b200.quantization.quantize_.common",
    "ao_b200.quantization.quantize_.workflows",
    "ao_b200.float8.inference",
    "ao_b200.prototype.mx_formats",
    # names as written by the reference: resolved onto our modules
    "torchao.quantization",
    "torchao.prototype.mx_formats",
    "torchao.float8.inference",
}

_MODULE_ALIASES = {
    "torchao.quantization": "ao_b200.quantization",
    "torchao.prototype.mx_formats": "ao_b200.prototype.mx_formats",
    "torchao.float8.inference": "ao_b200.float8.inference",
}


def _encode(o: Any) -> Any:
    if isinstance(o, AOBaseConfig) or (dataclasses.is_dataclass(o) and not isinstance(o, type)):
        data = {}
        if dataclasses.is_dataclass(o):
            for f in dataclasses.fields(o):
                if f.name == "version":
                    continue
                data[f.name] = _encode(getattr(o, f.name))
        return {"_type": type(o).__name__, "_version": getattr(o, "version", 1), "_data": data,
                "_module": type(o).__module__}
    if isinstance(o, tuple) and hasattr(o, "_fields"):  # NamedTuple (e.g. Float8MMConfig)
        return {"_type": type(o).__name__, "_version": 1, "_data": {k: _encode(v) for k, v in o._asdict().items()},
                "_module": type(o).__module__}
    if isinstance(o, enum.Enum):
        return {"_type": type(o).__name__, "_data": o.name, "_module": type(o).__module__, "_enum": True}
    if isinstance(o, torch.dtype):
        return {"_type": "torch.dtype", "_data": str(o).split(".")[-1]}
    if isinstance(o, (list, tuple)):
        return [_encode(v) for v in o]
    if isinstance(o, dict):
        return {k: _encode(v) for k, v in o.items()}
    if isinstance(o, torch.Tensor):
        return {"_type": "torch.Tensor", "_data": o.tolist(), "_dtype": str(o.dtype).split(".")[-1]}
    return o


class ConfigJSONEncoder(json.JSONEncoder):
    def default(self, o):
        enc = _encode(o)
        if enc is o:
            return super().default(o)
        return enc


def config_to_dict(config: AOBaseConfig) -> Dict[str, Any]:
    if not isinstance(config, AOBaseConfig):
        raise TypeError(f"expected an AOBaseConfig, got {type(config)}")
    return json.loads(json.dumps(_encode(config)))


def _resolve(module: str, name: str):
    base = module
    for allowed in sorted(ALLOWED_AO_MODULES, key=len, reverse=True):
        if module == allowed or module.startswith(allowed + "."):
            break
    else:
        raise ValueError(f"refusing to import config type {name} from non-allowlisted module {module}")
    for src, dst in _MODULE_ALIASES.items():
        if base == src or base.startswith(src + "."):
            base = dst
            break
    for cand in (base, "ao_b200.quantization", "ao_b200.prototype.mx_formats", "ao_b200.float8.inference",
                 "ao_b200.quantization.quant_primitives", "ao_b200.quantization.quantize_.common"):
        try:
            mod = importlib.import_module(cand)
        except ImportError:
            continue
        if hasattr(mod, name):
            return getattr(mod, name)
    raise ValueError(f"unknown config type {name} (module {module})")


def _decode(o: Any) -> Any:
    if isinstance(o, list):
        return [_decode(v) for v in o]
    if isinstance(o, dict):
        if "_type" in o and "_data" in o:
            t = o["_type"]
            if t == "torch.dtype":
                return getattr(torch, o["_data"])
            if t == "torch.Tensor":
                return torch.tensor(o["_data"], dtype=getattr(torch, o.get("_dtype", "float32")))
            cls = _resolve(o.get("_module", "ao_b200.quantization"), t)
            if o.get("_enum") or (isinstance(cls, type) and issubclass(cls, enum.Enum)):
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
