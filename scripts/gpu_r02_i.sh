#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.log)
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for l in open('gpurun_out/bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        for k in ('value','ms_per_step','e2e','gpu_launches'):
            print(k, d.get(k))
        print({k:(round(v['ms']*1e3,1) if isinstance(v,dict) and 'ms' in v else v) for k,v in (d.get('roofline_kernels') or {}).items()})
    else: print(l.strip()[:300])
PY
