#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/kernel_cases.py --case k3 --iters 3 > gpurun_out/plain_k3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_k3 -s 5 -c 2 -f -o gpurun_out/k_k3 \
    python scripts/kernel_cases.py --case k3 --iters 3 > gpurun_out/ncu_k3.log 2>&1
cat gpurun_out/plain_k3.log
