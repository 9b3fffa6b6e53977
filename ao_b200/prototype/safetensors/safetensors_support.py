"""state_dict <-> (plain tensors, string metadata) for safetensors.

Same wire format as `torchao/prototype/safetensors/safetensors_support.py:15-201`: tensor attribute ``a`` of the
quantized parameter ``<module>.<param>`` is stored under the key ``<module>._<param>_<a>``; the metadata dict maps
each original tensor name to its JSON description and ``"tensor_names"`` to the JSON list of names.  Plain
``torch.Tensor`` entries keep their key.  `unflatten` accepts partial shards (sharded checkpoints): tensors whose
pieces are not all present yet are skipped and their pieces returned in the leftover dict.
"""
from __future__ import annotations

import json
from typing import Any, Dict, Tuple

import torch

from .safetensors_utils import ALLOWED_TENSORS_SUBCLASSES, TensorSubclassAttributeJSONEncoder, object_from_dict

__all__ = ["flatten_tensor_state_dict", "unflatten_tensor_state_dict"]


def _piece_prefix(tensor_name: str) -> str:
    module_fqn, param = tensor_name.rsplit(".", 1)
    return f"{module_fqn}._{param}_"


def flatten_tensor_state_dict(tensors_dict: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    flat: Dict[str, torch.Tensor] = {}
    metadata: Dict[str, str] = {}
    for name, tensor in tensors_dict.items():
        if type(tensor).__name__ in ALLOWED_TENSORS_SUBCLASSES:
            prefix = _piece_prefix(name)
            for attr in list(tensor.tensor_data_names) + list(getattr(tensor, "optional_tensor_data_names", None) or []):
                piece = getattr(tensor, attr)
                if piece is not None:
                    flat[prefix + attr] = piece.detach().clone() if isinstance(piece, torch.Tensor) else piece
            metadata[name] = json.dumps(tensor, cls=TensorSubclassAttributeJSONEncoder)
        elif type(tensor) is torch.Tensor:
            flat[name] = tensor
            metadata[name] = json.dumps({"_type": torch.Tensor.__name__})
        else:
            raise ValueError(f"Unsupported tensor type: {type(tensor)}")
    metadata["tensor_names"] = json.dumps(list(tensors_dict.keys()))
    return flat, metadata


def unflatten_tensor_state_dict(tensors_data_dict: Dict[str, Any], metadata: Dict[str, Any]):
    if "tensor_names" not in metadata:
        raise ValueError("No tensors found")
    rebuilt: Dict[str, torch.Tensor] = {}
    leftover = dict(tensors_data_dict)
    for name in json.loads(metadata["tensor_names"]):
        desc = json.loads(metadata.get(name))
        kind = desc.get("_type")
        if kind in ALLOWED_TENSORS_SUBCLASSES:
            prefix = _piece_prefix(name)
            pieces = {k[len(prefix):]: v for k, v in tensors_data_dict.items() if k.startswith(prefix)}
            expected = desc.get("_tensor_data_names")
            if len(pieces) != len(expected):
                continue  # the rest arrives with a later shard
            desc["_data"].update(pieces)
            rebuilt[name] = object_from_dict(desc)
            for attr in expected:
                del leftover[prefix + attr]
        elif kind == torch.Tensor.__name__:
            if name not in tensors_data_dict:
                continue
            rebuilt[name] = tensors_data_dict[name]
            del leftover[name]
        else:
            raise ValueError(f"Unsupported tensor type: {kind}")
    return rebuilt, leftover
