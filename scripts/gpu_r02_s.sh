#!/bin/bash
mkdir -p gpurun_out
for m in 1 0; do WF_AB_OLD=$m timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn | tee -a gpurun_out/upsample_probe.log; done
(timeout 900 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "upsample" > gpurun_out/pytest_up.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_up.log)
tail -3 gpurun_out/pytest_up.log | cut -c1-200
