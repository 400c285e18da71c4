#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
for c in c4 dwconv convt head layernorm; do timeout 300 python scripts/kernel_cases.py --case $c --iters 10; done > gpurun_out/kernel_cases.log 2>&1
timeout 300 python scripts/module_times.py > gpurun_out/module_times.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/kernel_cases.log; grep -E "forward|block1.0|encoder1$|learnable_up[34]$|decoder1$|waveformer_encoder$" gpurun_out/module_times.log
