#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_inferer.py -m gpu -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_model.log)
tail -4 gpurun_out/pytest_model.log | cut -c1-250
for cfg in 42 51; do
if [ $cfg = 51 ]; then export WF_K3_ROLL_51=1; fi
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$cfg.log 2> gpurun_out/bench_$cfg.err; echo "bench exit $?" >> gpurun_out/bench_$cfg.log)
tail -3 gpurun_out/bench_$cfg.err
python - <<PY
import json
for l in open('gpurun_out/bench_$cfg.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('cfg $cfg', d.get('value'), d.get('ms_per_step'), d['e2e']['value'], d.get('gpu_launches'))
        print({k:(round(v['ms']*1e3,1) if isinstance(v,dict) and 'ms' in v else v) for k,v in (d.get('roofline_kernels') or {}).items() if 'k3' in k or 'attention' in k})
    else: print(l.strip()[:300])
PY
done
