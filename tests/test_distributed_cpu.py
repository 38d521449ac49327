"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: unit partitioning and the spanning-call
scatter/compute/gather, with the oracle injected as the compute function (the product default is the CUDA kernel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG_PY  # noqa: F401


def test_partition_units_properties():
    from aule.distributed import partition_units
    for units in (1, 2, 7, 8, 64, 256):
        for world in (1, 2, 3, 4, 8):
            parts = partition_units(units, world)
            assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == units
            sizes = [b - a for a, b in parts]
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_compute(q, k, v, causal, scale):
    from oracle import attention_oracle as orc
    o, _ = orc.attention_ref(q.numpy(), k.numpy(), v.numpy(), causal=causal, scale=scale, acc=np.float64)
    return torch.from_numpy(o.astype(np.float32))


def _worker(rank, world, port, shape, causal, ret):
    import sys
    for p in (ROOT, PKG_PY):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aule.distributed import flash_attention_spanning
    B, Hq, Hkv, Sq, Sk, D = shape
    q = k = v = None
    if rank == 0:
        g = torch.Generator().manual_seed(42)
        q = torch.randn(B, Hq, Sq, D, generator=g)
        k = torch.randn(B, Hkv, Sk, D, generator=g)
        v = torch.randn(B, Hkv, Sk, D, generator=g)
    out = flash_attention_spanning(q, k, v, causal=causal, src=0, device=torch.device("cpu"), compute=_oracle_compute)
    if rank == 0:
        full = _oracle_compute(q, k, v, causal, None)
        ret["err"] = float((out - full).abs().max())
        ret["shape"] = tuple(out.shape)
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,causal", [((2, 8, 2, 24, 24, 16), True),      # GQA, units=4 over 2 ranks
                                          ((1, 3, 3, 10, 17, 8), False),     # odd unit count, cross attention
                                          ((1, 4, 1, 12, 12, 8), True)])     # MQA: 1 unit, one rank idle
def test_spanning_call_gloo_world2(shape, causal):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, shape, causal, ret), nprocs=world, join=True)
    assert ret["shape"] == (shape[0], shape[1], shape[3], shape[5])
    assert ret["err"] < 1e-6


def test_shard_heads():
    from aule.distributed import shard_heads
    t = torch.arange(1 * 32 * 2 * 2).reshape(1, 32, 2, 2)
    parts = [shard_heads(t, 32, r, 8) for r in range(8)]
    assert all(p.shape[1] == 4 for p in parts)
    assert torch.equal(torch.cat(parts, dim=1), t)
    tq = torch.arange(2 * 32).reshape(2, 32, 1, 1)
    assert torch.equal(torch.cat([shard_heads(tq, 8, r, 4) for r in range(4)], dim=1), tq)   # GQA: 8 kv heads, group 4
