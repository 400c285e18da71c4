#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -m gpu -q -x -k "instance_norm" > gpurun_out/pytest_y.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_y.log)
tail -4 gpurun_out/pytest_y.log | cut -c1-300
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 5 --warm 2 2>&1 | tail -1
