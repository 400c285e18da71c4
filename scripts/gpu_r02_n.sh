#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "instance_norm or projection or pw_gelu" > gpurun_out/pytest_in.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_in.log)
tail -3 gpurun_out/pytest_in.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest_in.log || exit 1
timeout 300 python scripts/kernel_cases.py --case instnorm --iters 10 2>&1 | tee gpurun_out/instnorm_times.log
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_q.log 2> gpurun_out/bench_q.err; echo "bench exit $?" >> gpurun_out/bench_q.log)
python - <<'PY'
import json
for l in open('gpurun_out/bench_q.log'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    else: print(l.strip()[:300])
PY
