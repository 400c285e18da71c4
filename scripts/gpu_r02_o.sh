#!/bin/bash
# launch list of the forward at the bench's batching (6 windows) + one --set full capture of the kernels furthest from their HBM floor
mkdir -p gpurun_out
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b6.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward.log 2>&1
cat gpurun_out/plain_forward.log | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'upsample_cell|convT_k2s2|patch_embed|instnorm_apply_head_voxel|conv3d_c4' -c 14 -f -o gpurun_out/r02_k_floor \
    python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 0 > gpurun_out/ncu_floor.log 2>&1
tail -2 gpurun_out/ncu_floor.log
ls -la gpurun_out/*.ncu-rep
