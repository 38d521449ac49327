"""One backward call of config C/2 (for ncu captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "aule-attention_b200", "python"))
import aule
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(4, 32, 4096, 128, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
k = torch.randn(4, 8, 4096, 128, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
v = torch.randn(4, 8, 4096, 128, device="cuda", dtype=torch.bfloat16, generator=g).requires_grad_()
o = aule.flash_attention(q, k, v, causal=True)
do = torch.randn_like(o)
for _ in range(3):
    q.grad = k.grad = v.grad = None
    o.backward(do, retain_graph=True)
torch.cuda.synchronize()
