#!/bin/bash
mkdir -p gpurun_out
echo "== k3 cases, ring 4 / 2 tiles"; timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 2>&1 | tee gpurun_out/k3_times_42.log
echo "== k3 cases, ring 5 / 1 tile"; WF_K3_ROLL_51=1 timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 2>&1 | grep -v cudnn | tee gpurun_out/k3_times_51.log
(timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -5 gpurun_out/pytest.log | cut -c1-250
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches','parity','sw_batch_2','tta8','cpu_baseline','clocks'):
            print(k, d.get(k))
        print({k:(v.get('ms'), v.get('frac')) for k,v in (d.get('roofline_kernels') or {}).items() if 'k3' in k or 'attention' in k})
        print('train', d.get('train_step'))
    else: print(l.strip()[:300])
PY
