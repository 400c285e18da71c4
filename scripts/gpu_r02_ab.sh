#!/bin/bash
# same-box A/B of the kernels that have a predecessor (WF_AB_OLD=1): device-resident voxels/s, short bench without the extras
mkdir -p gpurun_out
nvidia-smi -q -d POWER | grep -i -E "limit|draw" | head -8
for i in 1 2 3; do
  for m in 1 0; do
    WF_AB_OLD=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB_OLD=$m', round(d['value']/1e6,2), 'M voxels/s', round(d['ms_per_step'],2), 'ms', d['clocks'])" | tee -a gpurun_out/ab.log
  done
done
