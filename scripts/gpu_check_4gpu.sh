#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 3 --warmup 3 --no-kernel-rooflines > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err; echo "bench n4 exit $?" >> gpurun_out/bench_n4.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_n4.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','n_gpus','e2e','strong','clocks'):
            print(k, d.get(k))
    elif 'exit' in l: print(l.strip())
PY
