"""Multi-GPU plumbing for the quantized-linear path: replicate the packed weights once, shard the batch.

The forward has no collective (each output row depends only on its input row and the replicated
weights).  One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is
used for exactly one thing: broadcasting rank 0's packed buffers at setup.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist

from ao_b200.utils import TorchAOBaseTensor


def packed_buffers(module: torch.nn.Module) -> List[Tuple[str, torch.Tensor]]:
    """Every plain tensor that makes up the quantized parameters of ``module`` (qdata, scales, ...), in a
    deterministic order, plus ordinary parameters/buffers.  Members of a fused group (``ao_b200.fusion``) hold views
    of the group's storage: the group's own (contiguous) buffers are listed once, under the first member's name."""
    from ao_b200.fusion import FusedLinearMember

    def flat(name, t, out):
        if isinstance(t, TorchAOBaseTensor):
            names, _ = t.__tensor_flatten__()
            for n in names:
                out.append((f"{name}.{n}", getattr(t, n)))
        elif t is not None:
            out.append((name, t))

    out, seen_groups, member_params = [], set(), set()
    for mod_name, mod in module.named_modules():
        if isinstance(mod, FusedLinearMember):
            g = mod._group
            for pn, _ in mod.named_parameters(recurse=False):
                member_params.add(f"{mod_name}.{pn}" if mod_name else pn)
            if id(g) not in seen_groups:
                seen_groups.add(id(g))
                prefix = f"{mod_name}.fused" if mod_name else "fused"
                flat(f"{prefix}.weight", g.weight, out)
                flat(f"{prefix}.bias", g.bias, out)
    for name, p in list(module.named_parameters()) + list(module.named_buffers()):
        if name in member_params:
            continue
        flat(name, p.data if isinstance(p, torch.nn.Parameter) else p, out)
    return out


def broadcast_packed_weights(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """Broadcast rank ``src``'s packed weights into every rank's (already allocated, same-shaped) buffers.
    Returns the number of bytes broadcast."""
    total = 0
    for _, t in packed_buffers(module):
        buf = t if t.dtype not in (torch.float8_e4m3fn, torch.float8_e8m0fnu) else t.view(torch.uint8)
        dist.broadcast(buf, src=src, group=group)
        total += t.numel() * t.element_size()
    return total


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [begin, end) of ``rank`` (earlier ranks get the remainder)."""
    base, rem = divmod(n_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    b, e = shard_rows(x.shape[0], rank, world)
    return x[b:e]


def gather_rows(y_local: torch.Tensor, n_rows: int, rank: int, world: int, group=None) -> torch.Tensor:
    """all_gather of row shards (reporting only; not part of the forward)."""
    sizes = [shard_rows(n_rows, r, world) for r in range(world)]
    mx = max(e - b for b, e in sizes)
    pad = torch.zeros(mx, *y_local.shape[1:], dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: e - b] for o, (b, e) in zip(outs, sizes)], dim=0)
