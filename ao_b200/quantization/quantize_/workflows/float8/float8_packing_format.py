from enum import Enum


class Float8PackingFormat(str, Enum):
    """reference: workflows/float8/float8_packing_format.py; only PLAIN has kernels here."""

    PLAIN = "plain"
    SPARSE_CUTLASS = "sparse_cutlass"
    OPAQUE = "opaque"
