#!/bin/bash
# round 2, pass C: GPU test suite, smoke, bench N=1
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -25 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches','parity','tta8','train_step','cpu_baseline','clocks'):
            print(k, d.get(k))
        print({k:(v['ms'],v['frac']) for k,v in d['roofline_kernels'].items()})
    else: print(l.strip()[:300])
PY
