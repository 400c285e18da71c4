#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "k3_c48 or k3_c96" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log)
tail -12 gpurun_out/pytest_k3.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest_k3.log || exit 1
timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 > gpurun_out/k3_times.log 2>&1
cat gpurun_out/k3_times.log
timeout 200 python scripts/k3_stage_clocks.py 2>&1 | tee gpurun_out/k3_stage_clocks.log
