#!/bin/bash
# launch lists of the window forward at 2 and 6 windows per forward (each ncu run directly follows a plain run of the same command)
mkdir -p gpurun_out
for b in 2 6; do
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward_b$b.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b$b.csv \
    python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward_b$b.log 2>&1
cat gpurun_out/plain_forward_b$b.log | tail -2
done
timeout 200 python scripts/k3_stage_clocks.py 2>&1 | tee gpurun_out/k3_stage_clocks.log
timeout 200 python scripts/k3_clock_check.py 2>&1 | cut -c1-330 | tee gpurun_out/k3_clock_check.log
