#!/bin/bash
mkdir -p gpurun_out
for c in c4 dwconv convt head; do
  timeout 300 python scripts/kernel_cases.py --case $c > gpurun_out/plain_$c.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv3d_c4|dwconv3d|layernorm|convT|instnorm_apply_head|attn_core|linear_tc|dwt_|idwt_' -s 2 -c 2 -f -o gpurun_out/k_$c \
      python scripts/kernel_cases.py --case $c > gpurun_out/ncu_$c.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
