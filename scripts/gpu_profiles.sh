#!/bin/bash
# ncu evidence for profiles/: every ncu run directly follows a plain run of the same command line.
mkdir -p gpurun_out
for c in haar c4 dwconv layernorm convt head attn; do
  timeout 300 python scripts/kernel_cases.py --case $c > gpurun_out/plain_$c.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv3d_c4|dwconv3d|layernorm|convT|instnorm_apply_head|attn_core|linear_tc|dwt_|idwt_' -s 2 -c 2 -f -o gpurun_out/k_$c \
      python scripts/kernel_cases.py --case $c > gpurun_out/ncu_$c.log 2>&1
  cat gpurun_out/plain_$c.log
done
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_forward_bf16.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward.log 2>&1
cat gpurun_out/plain_forward.log
