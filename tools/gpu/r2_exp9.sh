#!/bin/bash
# 2 GPUs: bench.py under torchrun (weak scaling line + config D strong scaling + native spanning call), distributed checks
set -u
OUT=gpurun_out/r2_exp9; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench2.json 2> $OUT/bench2.err; echo "bench2 rc=$?"; tail -5 $OUT/bench2.err
python -c "
import json;d=json.load(open('$OUT/bench2.json'));print(json.dumps({k:d[k] for k in ('value','ms_per_step','clocks','e2e')},indent=1)[:2500]);s=d['secondary'];print(json.dumps({k:s[k] for k in ('config_D_head_sharded','spanning_call')},indent=1)[:4000])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $OUT/bench1.json 2> $OUT/bench1.err; echo "bench1 rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench1.json'));print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['copy_ceiling'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_spanning_nccl.py > $OUT/spanning_nccl.log 2>&1; echo "nccl spanning rc=$?"; tail -3 $OUT/spanning_nccl.log
