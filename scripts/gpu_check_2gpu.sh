#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x -k "nccl or two_gpu or 2gpu or sharded" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_2gpu.log)
tail -3 gpurun_out/pytest_2gpu.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?" >> gpurun_out/bench_n2.log
tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','n_gpus','scaling','e2e','strong','clocks'):
            print(k, d.get(k))
        print(d['config']['workload'])
    else: print(l.strip()[:300])
PY
