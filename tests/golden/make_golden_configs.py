"""Golden fixture for config JSON (what HF TorchAoConfig stores in config.json): config_to_dict of the north-star
configs by the REFERENCE.   PYTHONPATH=/root/reference python tests/golden/make_golden_configs.py -> config_json.json"""
import json
import os
import sys
from collections import OrderedDict

REF = os.environ.get("AO_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from torchao.core.config import config_to_dict
    from torchao.prototype.mx_formats import (MXDynamicActivationMXWeightConfig, NVFP4DynamicActivationNVFP4WeightConfig,
                                              NVFP4WeightOnlyConfig)
    from torchao.quantization import (Float8DynamicActivationFloat8WeightConfig, FqnToConfig, Int4WeightOnlyConfig,
                                      Int8DynamicActivationInt8WeightConfig, PerRow)

    cfgs = OrderedDict()
    cfgs["int4_tile_g32"] = Int4WeightOnlyConfig(group_size=32, int4_packing_format="tile_packed_to_4d")
    cfgs["int4_tile_g128_hqq"] = Int4WeightOnlyConfig(group_size=128, int4_packing_format="tile_packed_to_4d",
                                                      int4_choose_qparams_algorithm="hqq")
    cfgs["int8_dyn"] = Int8DynamicActivationInt8WeightConfig()
    cfgs["fp8_rowwise"] = Float8DynamicActivationFloat8WeightConfig(granularity=PerRow())
    cfgs["fp8_default"] = Float8DynamicActivationFloat8WeightConfig()
    cfgs["mxfp8"] = MXDynamicActivationMXWeightConfig()
    cfgs["nvfp4_dyn"] = NVFP4DynamicActivationNVFP4WeightConfig()
    cfgs["nvfp4_wo"] = NVFP4WeightOnlyConfig(use_dynamic_per_tensor_scale=False)
    cfgs["fqn"] = FqnToConfig(OrderedDict([("re:.*q_proj", cfgs["int4_tile_g32"]), ("lm_head", None),
                                           ("_default", cfgs["int8_dyn"])]))
    out = {k: config_to_dict(v) for k, v in cfgs.items()}
    with open(os.path.join(HERE, "config_json.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out["fp8_rowwise"])[:600])


if __name__ == "__main__":
    main()
