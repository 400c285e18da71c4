#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "k3_c48 or k3_c96" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log)
tail -4 gpurun_out/pytest_k3.log | cut -c1-200
timeout 300 python scripts/k3_trace.py 2>&1 | grep "^run" | cut -c1-330
timeout 300 python scripts/k3_clock_check.py 2>&1 | grep -v "^  cta" | cut -c1-330
