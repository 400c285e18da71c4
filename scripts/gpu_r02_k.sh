#!/bin/bash
# launch lists of the window forward at 2 and 6 windows per forward (each ncu run directly follows a plain run of the same command),
# then the whole GPU suite and the bench line
mkdir -p gpurun_out
for b in 2 6; do
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward_b$b.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b$b.csv \
    python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward_b$b.log 2>&1
cat gpurun_out/plain_forward_b$b.log | tail -1
done
(timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -3 gpurun_out/pytest.log | cut -c1-250
(timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -2 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches','parity','sw_batch_2','tta8','cpu_baseline','clocks'):
            print(k, d.get(k))
        print({k:(v.get('ms'), v.get('frac')) for k,v in (d.get('roofline_kernels') or {}).items() if 'k3' in k or 'attention' in k or 'roundtrip_ncdhw_bf16' in k})
        print('train', d.get('train_step'))
    else: print(l.strip()[:300])
PY
(timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/bench_ref.log); tail -c 700 gpurun_out/bench_ref.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log); tail -2 gpurun_out/smoke.log
