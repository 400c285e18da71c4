#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "convt or conv_transpose or convT or transpose" > gpurun_out/pytest_z.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_z.log)
tail -3 gpurun_out/pytest_z.log | cut -c1-300
timeout 300 python scripts/kernel_cases.py --case convt --iters 10 2>&1 | grep -v Warn | tee gpurun_out/convt_times.log
