#!/bin/bash
# --set full capture of the last session's replacement kernels inside a six-window forward (after a plain run of the same command)
mkdir -p gpurun_out
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 0 > gpurun_out/plain_forward_v.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'instnorm_apply_reg|convT_k2s2_persist|head_voxel2|patch_embed_k2s2_c4_x4|upsample_cell_ring|conv3d_c4' -c 24 -f -o gpurun_out/r02_k_final \
    python scripts/profile_forward.py --dtype bf16 --batch 6 --no-profiler --iters 1 --warm 0 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log; tail -1 gpurun_out/plain_forward_v.log
