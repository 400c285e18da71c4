#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_attention.py -q > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_attn.log)
timeout 300 python scripts/attn_bench.py --fmt fp16 > gpurun_out/attn_bench.log 2>&1
timeout 300 python scripts/attn_bench.py --fmt fp16 --only s1.L1 --iters 3 > gpurun_out/plain_attn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn_core|linear_tc' -s 9 -c 3 -f -o gpurun_out/attn_s1L1 \
    python scripts/attn_bench.py --fmt fp16 --only s1.L1 --iters 3 > gpurun_out/ncu_attn.log 2>&1
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_forward_bf16.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward.log 2>&1
tail -4 gpurun_out/pytest_attn.log; cat gpurun_out/attn_bench.log
