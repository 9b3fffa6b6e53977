from enum import Enum


class Int4ChooseQParamsAlgorithm(str, Enum):
    """How scale / zero_point are chosen (reference: workflows/int4/int4_choose_qparams_algorithm.py)."""

    TINYGEMM = "tinygemm"
    HQQ = "hqq"
