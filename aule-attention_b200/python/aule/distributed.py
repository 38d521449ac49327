"""
Multi-GPU plumbing for the attention hot path.

The path shards with NO data-path collective: every (batch, kv-head) "unit" -- one KV head with the
Hq/Hkv query heads that share it -- is an independent problem.  One process per GPU
(`torch.distributed`, NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests):

  * resident inputs (the normal data-parallel case): each rank simply calls `aule.flash_attention`
    on its own shard; nothing here is needed.
  * spanning call: one rank holds q, k, v.  `flash_attention_spanning` scatters contiguous unit slabs
    with batched point-to-point sends, every rank runs the kernel on its slab, and O is gathered
    back on the source rank.  The reference has no multi-GPU code at all; this is the
    "scatter inputs / gather outputs" helper BASELINE.json's north_star allows, and the only place
    NCCL is used.

`compute` is injectable so the world_size-2 gloo tests can exercise the plumbing on CPU with the
oracle standing in for the kernel; the default is the CUDA kernel (`aule.flash_attention`).
"""
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def partition_units(units: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced split of `units` = B*Hkv (batch x kv-head) units over `world` ranks.
    The q-heads of a KV group stay together, so no K/V duplication."""
    base, extra = divmod(units, world)
    out, u = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((u, u + n))
        u += n
    return out


def _unit_views(q, k, v):
    B, Hq, Sq, D = q.shape
    Hkv, Sk = k.shape[1], k.shape[2]
    g = Hq // Hkv
    return q.reshape(B * Hkv, g, Sq, D), k.reshape(B * Hkv, 1, Sk, D), v.reshape(B * Hkv, 1, Sk, D)


def _default_compute(q, k, v, causal, scale):
    from . import flash_attention
    return flash_attention(q, k, v, causal=causal, scale=scale)


def flash_attention_spanning(q: Optional[torch.Tensor], k: Optional[torch.Tensor], v: Optional[torch.Tensor],
                             causal: bool = True, scale: Optional[float] = None, src: int = 0, group=None,
                             device: Optional[torch.device] = None,
                             compute: Callable = _default_compute) -> Optional[torch.Tensor]:
    """q: [B,Hq,Sq,D], k/v: [B,Hkv,Sk,D] valid on rank `src` (other ranks pass None). Returns O on `src`,
    None elsewhere. Scatter and gather are grouped point-to-point transfers of contiguous slabs."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    meta = [None]
    if rank == src:
        if q.shape[1] % k.shape[1] != 0:
            raise ValueError(f"heads_q ({q.shape[1]}) must be divisible by heads_kv ({k.shape[1]}) for GQA")
        meta = [(tuple(q.shape), tuple(k.shape), q.dtype)]
        device = q.device
    dist.broadcast_object_list(meta, src=src, group=group)
    (B, Hq, Sq, D), (_, Hkv, Sk, _), dtype = meta[0]
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    g = Hq // Hkv
    parts = partition_units(B * Hkv, world)
    u0, u1 = parts[rank]
    nu = u1 - u0

    if rank == src:
        qu, ku, vu = _unit_views(q.contiguous(), k.contiguous(), v.contiguous())
        ops = []
        for r, (a, b) in enumerate(parts):
            if r == src or b == a:
                continue
            ops += [dist.P2POp(dist.isend, t[a:b].contiguous(), r, group) for t in (qu, ku, vu)]
        reqs = dist.batch_isend_irecv(ops) if ops else []
        ql, kl, vl = qu[u0:u1], ku[u0:u1], vu[u0:u1]
    else:
        ql = torch.empty((nu, g, Sq, D), dtype=dtype, device=device)
        kl = torch.empty((nu, 1, Sk, D), dtype=dtype, device=device)
        vl = torch.empty((nu, 1, Sk, D), dtype=dtype, device=device)
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.irecv, t, src, group) for t in (ql, kl, vl)]) if nu else []
    for rq in reqs:
        rq.wait()

    ol = compute(ql, kl, vl, causal, scale) if nu else torch.empty((0, g, Sq, D), dtype=dtype, device=device)

    if rank == src:
        out = torch.empty((B * Hkv, g, Sq, D), dtype=ol.dtype, device=device)
        out[u0:u1] = ol
        ops = [dist.P2POp(dist.irecv, out[a:b], r, group) for r, (a, b) in enumerate(parts) if r != src and b > a]
        for rq in (dist.batch_isend_irecv(ops) if ops else []):
            rq.wait()
        return out.reshape(B, Hq, Sq, D)
    if nu:
        for rq in dist.batch_isend_irecv([dist.P2POp(dist.isend, ol.contiguous(), src, group)]):
            rq.wait()
    return None


def shard_heads(t: torch.Tensor, hkv_total: int, rank: int, world: int) -> torch.Tensor:
    """Resident-input helper: the contiguous head range of a [B,H,S,D] tensor owned by `rank` when the
    KV heads are split evenly (config D: [1,32,32768,128] -> 4 heads per GPU at world=8)."""
    B, H = t.shape[0], t.shape[1]
    g = H // hkv_total
    per = hkv_total // world
    return t[:, rank * per * g:(rank + 1) * per * g]
