#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/pytest_q.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_q.log)
tail -4 gpurun_out/pytest_q.log | cut -c1-300
timeout 300 python scripts/kernel_cases.py --case head --iters 10 2>&1 | grep -v Warn | tee gpurun_out/head_times.log
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 5 --warm 2 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b6.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 0 > gpurun_out/ncu_forward.log 2>&1
