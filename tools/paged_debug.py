import os, sys, torch
sys.path.insert(0, "aule-attention_b200/python"); sys.path.insert(0, ".")
import aule
from oracle import attention_oracle as orc
B, Hq, Hkv, D, bs, ctx = int(os.environ.get("PB", 16)), 32, 8, 128, 16, int(os.environ.get("PCTX", 2048))
g = torch.Generator().manual_seed(3)
mb = ctx // bs; nb = B * mb
q = torch.randn(B, Hq, D, generator=g).bfloat16().cuda()
kc = torch.randn(nb, bs, Hkv, D, generator=g).bfloat16().cuda(); vc = torch.randn(nb, bs, Hkv, D, generator=g).bfloat16().cuda()
bt = torch.randperm(nb, generator=g).reshape(B, mb).to(torch.int32).cuda()
cl = torch.full((B,), ctx, dtype=torch.int32).cuda()
for rep in range(3):
    out = aule.flash_attention_paged(q, kc, vc, bt, cl, max_context_len=ctx)
torch.cuda.synchronize()
print("ran", flush=True)
exp = orc.paged_decode_ref(q[:2].float().cpu().numpy(), kc.float().cpu().numpy(), vc.float().cpu().numpy(), bt[:2].cpu().numpy(), cl[:2].cpu().numpy())
print("err", orc.rel_err_to_scale(out[:2].float().cpu().numpy(), exp), flush=True)
