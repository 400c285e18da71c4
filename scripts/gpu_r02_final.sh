#!/bin/bash
# final check of the committed build: full GPU suite, smoke, same-box A/B (WF_AB_OLD=1: predecessors of convT / patch embedding / head / upsampling)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -3 gpurun_out/pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
rm -f gpurun_out/ab.log
for i in 1 2 3; do
  for m in 1 0; do
    WF_AB_OLD=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('AB_OLD=$m', round(d['value']/1e6,2), 'M voxels/s', round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value']/1e6,2), d['clocks'])" | tee -a gpurun_out/ab.log
  done
done
