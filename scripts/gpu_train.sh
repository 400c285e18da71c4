#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_model.py -q -x -k "gradients" > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log)
tail -15 gpurun_out/pytest_train.log | cut -c1-220
timeout 600 python scripts/train_step.py > gpurun_out/train_fp32.log 2>&1; tail -3 gpurun_out/train_fp32.log | cut -c1-400
timeout 600 python scripts/train_step.py --amp > gpurun_out/train_amp.log 2>&1; tail -3 gpurun_out/train_amp.log | cut -c1-400
