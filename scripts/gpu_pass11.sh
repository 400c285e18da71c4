#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_inferer.py -q > gpurun_out/pytest_inf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_inf.log)
(timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
(timeout 900 python bench.py --no-cuda-graph --no-cpu-baseline --no-kernel-rooflines > gpurun_out/bench_eager.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_eager.log)
tail -5 gpurun_out/pytest_inf.log; tail -c 300 gpurun_out/bench.log;  tail -c 300 gpurun_out/bench_eager.log
