#!/bin/bash
# windows per forward: 6 (default) vs 9 vs 18 - device-resident voxels/s on the same box
mkdir -p gpurun_out
for b in 6 9 18 6 9; do
  timeout 400 python bench.py --steps 8 --warmup 3 --sw-batch $b --no-cpu-baseline --no-kernel-rooflines --no-extras 2>gpurun_out/swb_$b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sw_batch=$b', round(d['value']/1e6,2), 'M voxels/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']/1e6,2), d['clocks']['sm_mhz'])" | tee -a gpurun_out/swb.log
done
tail -2 gpurun_out/swb_18.err
