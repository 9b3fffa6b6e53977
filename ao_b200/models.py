"""Synthetic Llama linear stacks used by bench.py / smoke(): ONLY the quantized nn.Linear forward
(the hot path), chained so every GEMM depends on the previous one.  Attention, norms, RoPE,
embeddings and lm_head are not part of the path and are not modelled.

Per layer (Llama-3-8B: hidden 4096, inter 14336, kv 1024):
    q = q_proj(x); k = k_proj(x); v = v_proj(x); o = o_proj(q)
    g = gate_proj(o); u = up_proj(o); x = down_proj(g)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.nn as nn


@dataclass(frozen=True)
class LlamaShape:
    name: str
    hidden: int
    inter: int
    kv: int
    layers: int

    def linears(self):
        h, i, kv = self.hidden, self.inter, self.kv
        return [("q_proj", h, h), ("k_proj", kv, h), ("v_proj", kv, h), ("o_proj", h, h),
                ("gate_proj", i, h), ("up_proj", i, h), ("down_proj", h, i)]

    def params_per_layer(self) -> int:
        return sum(n * k for _, n, k in self.linears())


LLAMA3_8B = LlamaShape("llama-3-8b", 4096, 14336, 1024, 32)
LLAMA3_70B = LlamaShape("llama-3-70b", 8192, 28672, 1024, 80)


class LlamaLinearLayer(nn.Module):
    def __init__(self, shape: LlamaShape, device, dtype=torch.bfloat16):
        super().__init__()
        for name, n, k in shape.linears():
            setattr(self, name, nn.Linear(k, n, bias=False, device=device, dtype=dtype))

    def forward(self, x):
        q = self.q_proj(x)
        self.k_proj(x)
        self.v_proj(x)
        o = self.o_proj(q)
        g = self.gate_proj(o)
        self.up_proj(o)
        return self.down_proj(g)


class LlamaLinearStack(nn.Module):
    """``layers`` x 7 linears with random-init weights of the real shapes (no checkpoints offline)."""

    def __init__(self, shape: LlamaShape, layers: int = None, device="cuda", seed: int = 0, init_scale: float = 0.02):
        super().__init__()
        self.shape = shape
        n = shape.layers if layers is None else layers
        gen = torch.Generator(device=device).manual_seed(seed)
        self.layers = nn.ModuleList()
        for _ in range(n):
            layer = LlamaLinearLayer(shape, device)
            with torch.no_grad():
                for p in layer.parameters():
                    p.copy_((torch.randn(p.shape, device=device, generator=gen) * init_scale).to(p.dtype))
            self.layers.append(layer)

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    def linear_modules(self) -> List[nn.Linear]:
        return [m for m in self.modules() if isinstance(m, nn.Linear)]
