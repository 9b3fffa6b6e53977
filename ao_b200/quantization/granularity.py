"""Quantization granularities (same names/fields as torchao/quantization/granularity.py:12-145)."""
from dataclasses import dataclass


@dataclass(frozen=True)
class Granularity:
    """Base class: how many elements share one set of quantization parameters."""


@dataclass(frozen=True)
class PerTensor(Granularity):
    """One scale for the whole tensor."""


@dataclass(frozen=True)
class PerAxis(Granularity):
    """One scale per slice along ``axis`` (reduction over every other dim)."""

    axis: int


@dataclass(frozen=True)
class PerGroup(Granularity):
    """One scale per ``group_size`` consecutive elements of the last dim."""

    group_size: int


@dataclass(frozen=True)
class PerRow(Granularity):
    """One scale per row; ``dim`` is the dimension that is reduced away (default: last)."""

    dim: int = -1


@dataclass(frozen=True)
class PerToken(Granularity):
    """One scale per token (all leading dims kept, last dim reduced)."""


@dataclass(frozen=True)
class PerBlock(Granularity):
    """One scale per block of ``block_size`` (tuple, one entry per dim)."""

    block_size: tuple
