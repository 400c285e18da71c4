#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "patch_embed" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_new.log)
tail -8 gpurun_out/pytest_new.log
grep -q "pytest exit 0" gpurun_out/pytest_new.log || exit 1
(timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -6 gpurun_out/pytest.log
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches','parity'):
            print(k, d.get(k))
    else: print(l.strip()[:300])
PY
