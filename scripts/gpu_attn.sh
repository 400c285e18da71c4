#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_attention_tc.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_attn.log)
tail -3 gpurun_out/pytest_attn.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest_attn.log || exit 1
timeout 300 python scripts/kernel_cases.py --case attn --iters 20 2>&1 | tee gpurun_out/attn_times.log
