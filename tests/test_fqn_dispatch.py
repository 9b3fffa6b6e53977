"""FqnToConfig dispatch order (exact parameter fqn > exact module fqn > parameter regex > module regex > `_default`,
`None` = skip) pinned to the reference: tests/golden/make_golden_fqn.py ran torchao's quantize_ with a tagging config
on CPU and recorded which config reached which module / parameter; the same cases must give the same outcome here."""
import json
import os
from collections import OrderedDict
from dataclasses import dataclass

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def build_model():
    return torch.nn.ModuleDict(OrderedDict(
        attn=torch.nn.ModuleDict(OrderedDict(q_proj=torch.nn.Linear(8, 8), k_proj=torch.nn.Linear(8, 4))),
        mlp=torch.nn.ModuleDict(OrderedDict(up=torch.nn.Linear(8, 16), act=torch.nn.ReLU(), down=torch.nn.Linear(16, 8))),
        norm=torch.nn.LayerNorm(8), head=torch.nn.Linear(8, 32, bias=False)))


with open(os.path.join(HERE, "golden", "fqn_dispatch.json")) as _f:
    GOLD = json.load(_f)


@pytest.fixture(scope="module")
def tag_config():
    import ao_b200  # noqa: F401
    from ao_b200.core.config import AOBaseConfig
    from ao_b200.quantization.transform_module import register_quantize_module_handler

    @dataclass
    class TagConfig(AOBaseConfig):
        tag: str = ""

    @register_quantize_module_handler(TagConfig)
    def _tag(module, config, *, parameter_name="weight"):
        module.__dict__.setdefault("_tags", []).append([config.tag, parameter_name])
        return module

    return TagConfig


@pytest.mark.parametrize("case", sorted(GOLD["cases"]))
def test_fqn_to_config_dispatch_matches_reference(case, tag_config):
    from ao_b200.quantization import FqnToConfig, quantize_

    model = build_model()
    cfg = FqnToConfig(OrderedDict((k, tag_config(v) if v is not None else None) for k, v in GOLD["cases"][case]))
    quantize_(model, cfg, filter_fn=None)
    got = {fqn: m.__dict__["_tags"] for fqn, m in model.named_modules() if "_tags" in m.__dict__}
    assert got == GOLD["expected"][case]
