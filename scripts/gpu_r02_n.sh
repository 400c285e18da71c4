#!/bin/bash
# round 2: multi-GPU bench, N = $1 (torchrun, one process per GPU)
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${2:-3} --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench exit $?" >> gpurun_out/bench_n$N.log)
grep -c "ncclReduce\|Reduce:" gpurun_out/bench_n$N.err gpurun_out/bench_n$N.log | head
grep -m3 "opCount.*Reduce\|Reduce: opCount" gpurun_out/bench_n$N.log gpurun_out/bench_n$N.err | cut -c1-300
grep -v "NCCL INFO" gpurun_out/bench_n$N.err | tail -15
python - <<PY
import json
for l in open('gpurun_out/bench_n$N.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('n_gpus','value','ms_per_step','scaling','e2e','gpu_launches','strong','clocks'):
            print(k, json.dumps(d.get(k)))
        print(d['config']['workload'])
    elif 'NCCL' not in l: print(l.strip()[:300])
PY
