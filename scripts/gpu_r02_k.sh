#!/bin/bash
# launch list of the window forward at 6 windows per forward (the ncu run directly follows a plain run of the same command)
mkdir -p gpurun_out
for b in 6; do
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward_b$b.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b$b.csv \
    python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward_b$b.log 2>&1
cat gpurun_out/plain_forward_b$b.log | tail -2
done
