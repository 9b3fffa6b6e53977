from enum import Enum


class Int4PackingFormat(str, Enum):
    """Packing formats of int4 weights (reference: workflows/int4/int4_packing_format.py).
    Only TILE_PACKED_TO_4D (the tinygemm layout, BASELINE's format) has kernels in this engine;
    the others need external libraries in the reference too (mslk) and raise here."""

    PLAIN = "plain"
    PRESHUFFLED = "preshuffled"
    PLAIN_INT32 = "plain_int32"
    TILE_PACKED_TO_4D = "tile_packed_to_4d"
