#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "instance_norm or c4 or shortcut" > gpurun_out/pytest_in.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_in.log)
tail -5 gpurun_out/pytest_in.log | cut -c1-300
timeout 300 python scripts/apply_stream_probe.py 2>&1 | grep -v Warning | tee gpurun_out/apply_probe.log
grep -q "pytest exit 0" gpurun_out/pytest_in.log || exit 1
timeout 600 python bench.py > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','parity','gpu_launches') if k in d})
P
