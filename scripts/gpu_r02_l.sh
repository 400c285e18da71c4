#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "pw_gelu or projection" > gpurun_out/pytest_pw.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_pw.log)
tail -12 gpurun_out/pytest_pw.log | cut -c1-220
grep -q "pytest exit 0" gpurun_out/pytest_pw.log || exit 1
(timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_inferer.py -m gpu -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_model.log)
tail -4 gpurun_out/pytest_model.log | cut -c1-250
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_q.log 2> gpurun_out/bench_q.err; echo "bench exit $?" >> gpurun_out/bench_q.log)
python - <<'PY'
import json
for l in open('gpurun_out/bench_q.log'):
    if l.startswith('{'):
        d=json.loads(l); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])
    else: print(l.strip()[:300])
PY
