#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --rows 80 > gpurun_out/prof_bf16.log 2>&1
(timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
grep -E "forward batch|Self CUDA time total|Self CPU time total" gpurun_out/prof_bf16.log; tail -c 400 gpurun_out/bench.log
