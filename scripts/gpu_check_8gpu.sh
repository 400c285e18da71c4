#!/bin/bash
# 8-GPU box: sharded-inferer check (one volume over 8 ranks + NCCL reduce), bench at N = 4 and N = 8
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/mgpu_check.py > gpurun_out/mgpu_check8.log 2>&1; echo "mgpu_check exit $?" >> gpurun_out/mgpu_check8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 --no-kernel-rooflines > gpurun_out/bench_n4.log 2>&1; echo "bench n4 exit $?" >> gpurun_out/bench_n4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --no-kernel-rooflines > gpurun_out/bench_n8.log 2>&1; echo "bench n8 exit $?" >> gpurun_out/bench_n8.log
grep -v -i "warn" gpurun_out/mgpu_check8.log | tail -4; tail -c 200 gpurun_out/bench_n4.log; tail -c 200 gpurun_out/bench_n8.log
