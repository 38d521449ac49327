#!/bin/bash
# 8 GPUs: bench.py under torchrun exactly as the driver launches it (reference arm first, then ours), N=8 and N=4
set -u
OUT=gpurun_out/r2_exp14; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo8.txt 2>&1
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench$N.json 2> $OUT/bench$N.err; echo "bench$N rc=$?"; tail -3 $OUT/bench$N.err
  python - <<PY
import json
txt=open('$OUT/bench$N.json').read().strip().splitlines()
print('stdout lines:', len(txt))
d=json.loads(txt[-1])
print(json.dumps({k:d[k] for k in ('value','ms_per_step','clocks')}))
print(json.dumps(d['e2e'])[:900])
s=d['secondary']
print(json.dumps({k:s.get(k) for k in ('config_D_head_sharded','spanning_call')},indent=1)[:3000])
PY
done
