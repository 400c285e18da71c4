#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -5 gpurun_out/pytest.log | cut -c1-300
timeout 300 python scripts/kernel_cases.py --case convt --iters 10 2>&1 | grep -v Warn | tee gpurun_out/convt_times.log
grep -q "pytest exit 0" gpurun_out/pytest.log || exit 1
timeout 600 python bench.py > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench exit $?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_n.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}, d['e2e']['value'], d['parity']['argmax_agreement'])
P
