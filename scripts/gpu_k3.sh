#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_glue.py -q -k "k3_c48" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log)
timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 > gpurun_out/k3_times.log 2>&1
tail -12 gpurun_out/pytest_k3.log | cut -c1-200; cat gpurun_out/k3_times.log
