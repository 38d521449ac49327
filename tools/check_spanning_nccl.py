"""torchrun --nproc-per-node N tools/check_spanning_nccl.py : spanning call over NCCL on real GPUs vs single-GPU result."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
import aule  # noqa: E402
from aule.distributed import flash_attention_spanning  # noqa: E402

rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
q = k = v = None
if rank == 0:
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(2, 16, 1024, 128, device="cuda", dtype=torch.bfloat16, generator=g)
    k = torch.randn(2, 4, 1024, 128, device="cuda", dtype=torch.bfloat16, generator=g)
    v = torch.randn(2, 4, 1024, 128, device="cuda", dtype=torch.bfloat16, generator=g)
out = flash_attention_spanning(q, k, v, causal=True, src=0)
if rank == 0:
    ref = aule.flash_attention(q, k, v, causal=True)
    print("spanning == local:", torch.equal(out, ref), "world", dist.get_world_size(), flush=True)
    assert torch.equal(out, ref)
dist.barrier()
dist.destroy_process_group()
