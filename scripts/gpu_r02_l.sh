#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_inferer.py -m gpu -q -x > gpurun_out/pytest_inf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_inf.log)
tail -3 gpurun_out/pytest_inf.log | cut -c1-200
timeout 300 python scripts/sw_overhead.py 2>&1 | grep -v "Warn\|warn" | tee gpurun_out/sw_overhead.log
