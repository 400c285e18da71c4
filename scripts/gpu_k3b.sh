#!/bin/bash
mkdir -p gpurun_out
for cfg in 42 51; do
  if [ $cfg = 51 ]; then export WF_K3_ROLL_51=1; fi
  echo "== ring/stages $cfg"
  timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 2>&1 | grep -v cudnn | tee gpurun_out/k3_times_$cfg.log
  timeout 200 python scripts/k3_stage_clocks.py 2>&1 | grep -v "^block\|loader   waits\|issuer   waits\|epilogue waits" | tee gpurun_out/k3_stage_clocks_$cfg.log
done
