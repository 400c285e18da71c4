#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -3 gpurun_out/pytest.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest.log || exit 1
for i in 1 2 3; do
  for m in kernel moments; do
    WF_C4_SHORTCUT_STATS=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('shortcut stats=$m', round(d['value']/1e6,2), 'M voxels/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab2.log
  done
done
