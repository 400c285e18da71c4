#!/bin/bash
# 8-GPU box: bench at N = 8 (64 volumes sharded; one volume over 8 ranks + NCCL reduce; interleaved arm)
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --no-kernel-rooflines > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 exit $?" >> gpurun_out/bench_n8.log
grep -m3 "NVLS\|NCCL version" gpurun_out/bench_n8.log gpurun_out/bench_n8.err | cut -c1-200
grep -m3 "Reduce:" gpurun_out/bench_n8.log gpurun_out/bench_n8.err | cut -c1-220
python - <<'PY'
import json
for l in open('gpurun_out/bench_n8.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','n_gpus','e2e','strong','clocks'):
            print(k, d.get(k))
    elif 'exit' in l: print(l.strip())
PY
