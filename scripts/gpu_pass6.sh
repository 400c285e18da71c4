#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
timeout 300 python scripts/module_times.py > gpurun_out/module_times.log 2>&1
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --rows 60 > gpurun_out/prof_bf16.log 2>&1
(timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
timeout 600 python scripts/precision_probe.py > gpurun_out/precision.log 2>&1
tail -12 gpurun_out/pytest.log; grep -E "forward|block1.0|encoder1$|learnable_up[34]$|decoder1$|waveformer_encoder$" gpurun_out/module_times.log; grep -E "policy, attention fp16|weights" gpurun_out/precision.log; tail -c 1500 gpurun_out/bench.log
