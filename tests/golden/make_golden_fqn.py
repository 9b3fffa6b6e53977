"""Golden fixture for FqnToConfig dispatch: run the REFERENCE's quantize_(model, FqnToConfig(...)) on CPU with a
tagging config (its handler records which config reached which module/parameter instead of quantizing) and store the
outcome per case.   PYTHONPATH=/root/reference python tests/golden/make_golden_fqn.py -> fqn_dispatch.json"""
import json
import os
import sys
from collections import OrderedDict
from dataclasses import dataclass

import torch

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "exact_module": [("attn.q_proj", "A")],
    "exact_param_beats_module": [("attn.q_proj.weight", "P"), ("attn.q_proj", "M")],
    "regex_module": [("re:attn\\\\..*_proj", "R")],
    "regex_param": [("re:.*\\\\.weight", "W")],
    "first_regex_wins": [("re:attn\\\\..*", "first"), ("re:.*q_proj", "second")],
    "none_skips": [("attn.q_proj", None), ("re:attn\\\\..*_proj", "R")],
    "default_fallback": [("mlp.up", "U"), ("_default", "D")],
    "exact_beats_regex": [("re:.*", "R"), ("mlp.up", "E")],
    "param_none_then_regex": [("attn.q_proj.weight", None), ("re:.*\\\\.weight", "W")],
    "bias_param": [("attn.q_proj.bias", "B")],
    "exact_param_then_regex_on_the_other_params": [("attn.q_proj.weight", "A"), ("re:.*bias", "B")],
    "two_regexes_on_different_params": [("re:.*q_proj\\.weight", "W"), ("re:.*q_proj\\.bias", "B")],
}


def build_model():
    return torch.nn.ModuleDict(OrderedDict(
        attn=torch.nn.ModuleDict(OrderedDict(q_proj=torch.nn.Linear(8, 8), k_proj=torch.nn.Linear(8, 4))),
        mlp=torch.nn.ModuleDict(OrderedDict(up=torch.nn.Linear(8, 16), act=torch.nn.ReLU(), down=torch.nn.Linear(16, 8))),
        norm=torch.nn.LayerNorm(8), head=torch.nn.Linear(8, 32, bias=False)))


def main():
    from torchao.core.config import AOBaseConfig
    from torchao.quantization import FqnToConfig, quantize_
    from torchao.quantization.transform_module import register_quantize_module_handler

    @dataclass
    class TagConfig(AOBaseConfig):
        tag: str = ""

    @register_quantize_module_handler(TagConfig)
    def _tag(module, config, *, parameter_name="weight"):
        module.__dict__.setdefault("_tags", []).append([config.tag, parameter_name])
        return module

    out = {}
    for name, items in CASES.items():
        model = build_model()
        cfg = FqnToConfig(OrderedDict((k.replace("\\\\", "\\"), (TagConfig(v) if v is not None else None)) for k, v in items))
        quantize_(model, cfg, filter_fn=None)
        out[name] = {fqn: m.__dict__["_tags"] for fqn, m in model.named_modules() if "_tags" in m.__dict__}
    with open(os.path.join(HERE, "fqn_dispatch.json"), "w") as f:
        json.dump({"cases": {k: [[a.replace("\\\\", "\\"), b] for a, b in v] for k, v in CASES.items()}, "expected": out}, f,
                  indent=1, sort_keys=True)
    print(json.dumps(out)[:1500])


if __name__ == "__main__":
    main()
