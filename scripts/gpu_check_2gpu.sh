#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py > gpurun_out/mgpu_check.log 2>&1; echo "mgpu_check exit $?" >> gpurun_out/mgpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 exit $?" >> gpurun_out/bench_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "bench ref exit $?" >> gpurun_out/bench_ref_n2.log
grep -v Warning gpurun_out/mgpu_check.log | tail -5; tail -c 900 gpurun_out/bench_n2.log; tail -c 600 gpurun_out/bench_ref_n2.log
