#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "c4 or convt or conv_transpose or convT" > gpurun_out/pytest_c4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c4.log)
tail -3 gpurun_out/pytest_c4.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest_c4.log || exit 1
timeout 300 python scripts/kernel_cases.py --case c4 --iters 10 2>&1 | tee gpurun_out/c4_times.log
timeout 300 python scripts/kernel_cases.py --case convt --iters 10 2>&1 | tee gpurun_out/convt_times.log
