"""Safetensors wire format: what the reference's flatten_tensor_state_dict wrote (golden fixture
made by tests/golden/make_golden_safetensors.py from torchao 0.19) must be rebuilt by our unflatten, and flattening the
rebuilt state dict must give back the same keys, bytes and metadata (compared as parsed JSON)."""
import json
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def wire():
    with open(os.path.join(HERE, "golden", "safetensors_wire.json")) as f:
        doc = json.load(f)
    raw = np.load(os.path.join(HERE, "golden", "safetensors_wire.npz"))
    flat = {}
    for key, info in doc["tensors"].items():
        dt = getattr(torch, info["dtype"])
        flat[key] = torch.from_numpy(raw[key.replace(".", "__")].copy()).view(dt).reshape(info["shape"])
    return flat, doc["metadata"]


def _parsed(meta):
    return {k: json.loads(v) for k, v in meta.items()}


def test_reference_metadata_is_recognised(wire):
    from ao_b200.prototype.safetensors.safetensors_utils import is_metadata_torchao

    _, meta = wire
    assert is_metadata_torchao(meta)
    assert not is_metadata_torchao({})
    assert not is_metadata_torchao({"tensor_names": "not json"})
    assert not is_metadata_torchao({"tensor_names": json.dumps(["a.w"]), "a.w": json.dumps({"_type": "Evil"})})


def test_unflatten_then_flatten_round_trips_the_reference_bytes(wire):
    import ao_b200  # noqa: F401
    from ao_b200.prototype.mx_formats import MXTensor, NVFP4Tensor
    from ao_b200.prototype.safetensors.safetensors_support import flatten_tensor_state_dict, unflatten_tensor_state_dict
    from ao_b200.quantization import Float8Tensor, Int4TilePackedTo4dTensor, Int8Tensor, PerRow

    flat, meta = wire
    sd, leftover = unflatten_tensor_state_dict(flat, meta)
    assert leftover == {}
    kinds = {"m.int4.weight": Int4TilePackedTo4dTensor, "m.int8.weight": Int8Tensor, "m.fp8.weight": Float8Tensor,
             "m.mx.weight": MXTensor, "m.nvfp4.weight": NVFP4Tensor}
    for name, cls in kinds.items():
        assert type(sd[name]) is cls and tuple(sd[name].shape) == (32, 128)
    assert type(sd["m.int8.bias"]) is torch.Tensor
    # nested objects are rebuilt as objects, not dicts
    assert isinstance(sd["m.fp8.weight"].act_quant_kwargs.granularity, PerRow)
    assert sd["m.int8.weight"].act_quant_kwargs.mapping_type.name == "SYMMETRIC"
    assert sd["m.mx.weight"].elem_dtype is torch.float8_e4m3fn and sd["m.nvfp4.weight"].orig_dtype is torch.bfloat16
    assert sd["m.int4.weight"].block_size == [1, 32]

    flat2, meta2 = flatten_tensor_state_dict(sd)
    assert set(flat2) == set(flat)
    for k in flat:
        assert flat2[k].dtype == flat[k].dtype and flat2[k].shape == flat[k].shape
        assert torch.equal(flat2[k].reshape(-1).view(torch.uint8), flat[k].reshape(-1).view(torch.uint8)), k
    assert _parsed(meta2) == _parsed(meta)


def test_partial_shards_and_errors(wire):
    from ao_b200.prototype.safetensors.safetensors_support import flatten_tensor_state_dict, unflatten_tensor_state_dict

    flat, meta = wire
    # a shard that misses one piece of the fp8 weight: that tensor is left for a later call, its pieces are returned
    shard = {k: v for k, v in flat.items() if k != "m.fp8._weight_scale"}
    sd, leftover = unflatten_tensor_state_dict(shard, meta)
    assert "m.fp8.weight" not in sd and "m.fp8._weight_qdata" in leftover and "m.int8.weight" in sd
    with pytest.raises(ValueError):
        unflatten_tensor_state_dict(flat, {k: v for k, v in meta.items() if k != "tensor_names"})
    bad = dict(meta)
    bad["m.int8.bias"] = json.dumps({"_type": "SomethingElse"})
    with pytest.raises(ValueError):
        unflatten_tensor_state_dict(flat, bad)
    with pytest.raises(ValueError):
        flatten_tensor_state_dict({"x.w": torch.nn.Parameter(torch.zeros(2))})


def test_quantize_then_save_load_through_safetensors_file(tmp_path, wire):
    """End to end through the real file format when the safetensors package is importable."""
    st = pytest.importorskip("safetensors.torch")
    from ao_b200.prototype.safetensors.safetensors_support import flatten_tensor_state_dict, unflatten_tensor_state_dict

    flat, meta = wire
    sd, _ = unflatten_tensor_state_dict(flat, meta)
    flat2, meta2 = flatten_tensor_state_dict(sd)
    path = str(tmp_path / "model.safetensors")
    try:
        st.save_file({k: v.contiguous() for k, v in flat2.items()}, path, metadata=meta2)
    except Exception as e:  # e8m0 / fp8 dtypes need a recent safetensors
        pytest.skip(f"safetensors cannot store these dtypes here: {e}")
    from safetensors import safe_open

    with safe_open(path, framework="pt") as f:
        meta3 = f.metadata()
        flat3 = {k: f.get_tensor(k) for k in f.keys()}
    sd3, leftover = unflatten_tensor_state_dict(flat3, meta3)
    assert leftover == {} and set(sd3) == set(sd)
    assert torch.equal(sd3["m.int4.weight"].qdata, sd["m.int4.weight"].qdata)
