#!/bin/bash
# attention kernels: parity tests, per-geometry timings, ncu capture of the stage-1 level-1 case
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_attention.py tests/test_gpu_model.py -q > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_attn.log)
timeout 300 python scripts/kernel_cases.py --case attn --iters 20 > gpurun_out/plain_attn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_core|linear_tc' -s 6 -c 3 -f -o gpurun_out/k_attn \
    python scripts/kernel_cases.py --case attn --iters 20 > gpurun_out/ncu_attn.log 2>&1
tail -4 gpurun_out/pytest_attn.log; cat gpurun_out/plain_attn.log
