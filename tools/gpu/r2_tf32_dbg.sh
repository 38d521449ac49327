#!/bin/bash
mkdir -p gpurun_out/r2_tf32
timeout 300 python - > gpurun_out/r2_tf32/dbg.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, "aule-attention_b200/python")
from aule import cuda_flash
torch.manual_seed(0)
S, D = 128, 64
q = torch.randn(1,1,S,D,device="cuda"); k = torch.randn(1,1,S,D,device="cuda"); v = torch.randn(1,1,S,D,device="cuda")
def run(q,k,v,name):
    o,l = cuda_flash.forward_with_lse(q,k,v,causal=False,allow_tf32=True)
    e,le = cuda_flash.forward_with_lse(q,k,v,causal=False,allow_tf32=False)
    torch.cuda.synchronize()
    print(name, "out absmax", o.abs().max().item(), "exp absmax", e.abs().max().item(), "err", (o-e).abs().max().item(),
          "| lse err", (l-le).abs().max().item(), "lse[0:3]", l[0,0,:3].tolist(), le[0,0,:3].tolist())
    return o, e
run(q,k,v,"random")
o,e = run(q,k,torch.ones_like(v),"V=1")
print(" V=1 out row0[:8]", o[0,0,0,:8].tolist())
o,e = run(torch.zeros_like(q),k,v,"Q=0")
print(" Q=0 out row0[:8]", o[0,0,0,:8].tolist(), "exp", e[0,0,0,:8].tolist())
vv = torch.zeros_like(v); vv[0,0,:,0] = 1.0
o,e = run(q,k,vv,"V=e0"); print(" V=e0 out row0[:8]", o[0,0,0,:8].tolist())
vv = torch.zeros_like(v); vv[0,0,0,:] = torch.arange(D,device="cuda").float()
o,e = run(torch.zeros_like(q),k,vv,"Q=0,V=row0 arange"); print(" out row0[:8]", o[0,0,0,:8].tolist(), "exp", e[0,0,0,:8].tolist()); print(" out row0[32:40]", o[0,0,0,32:40].tolist(), "exp", e[0,0,0,32:40].tolist())
vv = torch.zeros_like(v); vv[0,0,5,:] = torch.arange(D,device="cuda").float()
o,e = run(torch.zeros_like(q),k,vv,"Q=0,V=row5 arange"); print(" out row0[:8]", o[0,0,0,:8].tolist(), "exp", e[0,0,0,:8].tolist())
PY
cat gpurun_out/r2_tf32/dbg.log | tail -20
