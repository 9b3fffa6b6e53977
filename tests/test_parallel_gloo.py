"""world_size-2 gloo test (CPU) of the multi-GPU host logic: one broadcast of packed weights, batch sharding."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ao_b200.parallel import broadcast_packed_weights, gather_rows, packed_buffers, shard_batch, shard_rows
        from ao_b200.quantization import Float8Tensor, Int8Tensor, PerRow
        from ao_b200.quantization.quantize_.workflows import QuantizeTensorToInt8Kwargs

        torch.manual_seed(100 + rank)  # different weights per rank before the broadcast
        m = torch.nn.Sequential(torch.nn.Linear(64, 32, bias=True, dtype=torch.bfloat16),
                                torch.nn.Linear(32, 16, bias=False, dtype=torch.bfloat16))
        m[0].weight = torch.nn.Parameter(Int8Tensor.from_hp(m[0].weight.data, PerRow(),
                                         act_quant_kwargs=QuantizeTensorToInt8Kwargs(granularity=PerRow())), requires_grad=False)
        m[1].weight = torch.nn.Parameter(Float8Tensor.from_hp(m[1].weight.data, granularity=PerRow()), requires_grad=False)
        names = [n for n, _ in packed_buffers(m)]
        assert names == ["0.weight.qdata", "0.weight.scale", "0.weight.zero_point", "0.bias", "1.weight.qdata", "1.weight.scale"]
        nbytes = broadcast_packed_weights(m, src=0)
        assert nbytes == 32 * 64 + 32 * 4 + 32 + 32 * 2 + 16 * 32 + 16 * 4
        # after the broadcast every rank holds rank 0's bytes
        flat = torch.cat([t.reshape(-1).view(torch.uint8).to(torch.float32) if t.element_size() == 1 else t.reshape(-1).float()
                          for _, t in packed_buffers(m)])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, flat)
        # batch sharding covers every row exactly once, also when the batch does not divide evenly
        for n in (7, 8, 1):
            rows = [shard_rows(n, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[-1][1] == n and all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
        x = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        y = gather_rows(shard_batch(x, rank, world) * 2, 7, rank, world)
        assert torch.equal(y, x * 2)
        ret[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        ret[rank] = f"{type(ex).__name__}: {ex}"
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    port = 29500 + (os.getpid() % 500)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
