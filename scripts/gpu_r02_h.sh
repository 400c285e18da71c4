#!/bin/bash
mkdir -p gpurun_out
for b in 2 3 6 9; do
  (timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-kernel-rooflines --sw-batch $b > gpurun_out/bench_swb$b.log 2> gpurun_out/bench_swb$b.err; echo "exit $?" >> gpurun_out/bench_swb$b.log)
  python - <<PY
import json
for l in open('gpurun_out/bench_swb$b.log'):
    if l.startswith('{'):
        d=json.loads(l); print('sw_batch $b', d['value'], d['ms_per_step'], d['e2e']['value'])
    else: print(l.strip()[:200])
PY
done
