#!/bin/bash
mkdir -p gpurun_out
for fork in 1 0; do for fused in 1 0; do
  WF_FORK=$fork WF_FFN_FUSED=$fused timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-kernel-rooflines --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fork=$fork fused=$fused ms_per_step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'])"
done; done
