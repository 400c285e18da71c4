#!/bin/bash
# round 2 ncu evidence for profiles/: every ncu run directly follows a plain run of the same command line.
mkdir -p gpurun_out
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b2.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward.log 2>&1
cat gpurun_out/plain_forward.log
for c in ffn attn k3; do
  timeout 300 python scripts/kernel_cases.py --case $c > gpurun_out/plain_$c.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ffn_front|ffn_back|attn_core_tc|linear_tc|conv3d_k3' -s 4 -c 4 -f -o gpurun_out/r02_k_$c \
      python scripts/kernel_cases.py --case $c > gpurun_out/ncu_$c.log 2>&1
  cat gpurun_out/plain_$c.log; tail -2 gpurun_out/ncu_$c.log
done
ls -la gpurun_out/*.ncu-rep
