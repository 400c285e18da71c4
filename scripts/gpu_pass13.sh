#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_inferer.py -q > gpurun_out/pytest_inf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_inf.log)
for sb in 2 3 6 9; do timeout 600 python bench.py --sw-batch $sb --no-cpu-baseline --no-kernel-rooflines --steps 4 > gpurun_out/bench_sb$sb.log 2>&1; done
tail -3 gpurun_out/pytest_inf.log
