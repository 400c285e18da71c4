#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_glue.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/pytest_t.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_t.log)
tail -3 gpurun_out/pytest_t.log | cut -c1-200
timeout 300 python scripts/kernel_cases.py --case ffn --iters 10 2>&1 | grep -v Warn | tee gpurun_out/ffn_times.log
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 5 --warm 2 2>&1 | tail -1
