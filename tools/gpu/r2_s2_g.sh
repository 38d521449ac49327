#!/bin/bash
set -u
OUT=gpurun_out/r2_s2_g; mkdir -p $OUT
cat > /tmp/one_f2.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "aule-attention_b200", "python"))
from aule import cuda_flash, ffi
lib = ffi.ensure_init()
lib.aule_set_kernel_path((1 << 26) | (8 << 18))
B, Hq, Hkv, S, D = 4, 32, 8, 4096, 128
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn(B, Hq, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
k = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
v = torch.randn(B, Hkv, S, D, device="cuda", dtype=torch.bfloat16, generator=g)
o, lse = cuda_flash.forward_with_lse(q, k, v, causal=True)
do = torch.randn_like(o)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(3):
    rc = lib.aule_attention_backward_dptr(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                                          B, Hq, Hkv, S, S, D, ffi.DTYPE_BF16, 0.0, 1, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aule_bwd_fused2 -s 1 -c 1 -o $OUT/bwd_fused2 -f python /tmp/one_f2.py > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
