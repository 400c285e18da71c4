#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
for c in c4 dwconv convt head; do timeout 300 python scripts/kernel_cases.py --case $c --iters 10; done > gpurun_out/kernel_cases.log 2>&1
timeout 300 python scripts/module_times.py > gpurun_out/module_times.log 2>&1
timeout 600 python scripts/precision_probe.py > gpurun_out/precision.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/kernel_cases.log; grep -E "forward|block1.0|encoder[1234]$|learnable_up[34]$|decoder1$|waveformer_encoder$" gpurun_out/module_times.log;  grep -E "policy, attention fp16|weights" gpurun_out/precision.log
