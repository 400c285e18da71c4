#!/bin/bash
# final round-2 checks: whole GPU suite, bench line, ncu --set full of the rolling-row convolution (after a plain run of the same command)
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -4 gpurun_out/pytest.log | cut -c1-250
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches','parity','sw_batch_2','tta8','cpu_baseline','clocks','roofline'):
            print(k, d.get(k))
        print({k:(v.get('ms'), v.get('frac')) for k,v in (d.get('roofline_kernels') or {}).items() if 'k3' in k or 'attention' in k})
        print('train', d.get('train_step'))
    else: print(l.strip()[:300])
PY
timeout 300 python scripts/kernel_cases.py --case k3 --iters 3 > gpurun_out/plain_k3roll.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3d_k3_c48_roll' -s 2 -c 2 -f -o gpurun_out/r02_k_k3roll \
    python scripts/kernel_cases.py --case k3 --iters 3 > gpurun_out/ncu_k3roll.log 2>&1
tail -2 gpurun_out/ncu_k3roll.log
ls -la gpurun_out/r02_k_k3roll.ncu-rep
