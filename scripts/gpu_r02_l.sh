#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "instance_norm or k3_c48" > gpurun_out/pytest_in.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_in.log)
tail -3 gpurun_out/pytest_in.log | cut -c1-200
timeout 300 python scripts/kernel_cases.py --case instnorm --iters 10 2>&1 | tee gpurun_out/instnorm_times.log
