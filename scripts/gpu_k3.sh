#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_glue.py -q -x -k "k3_c48 or k3_c96" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log)
tail -3 gpurun_out/pytest_k3.log | cut -c1-200
grep -q "pytest exit 0" gpurun_out/pytest_k3.log || exit 1
timeout 300 python scripts/kernel_cases.py --case k3 --iters 10 > gpurun_out/k3_times.log 2>&1
cat gpurun_out/k3_times.log
timeout 200 python scripts/k3_stage_clocks.py 2>&1 | tee gpurun_out/k3_stage_clocks.log | head -3
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv3d_k3_c48_roll' -s 2 -c 2 --csv --log-file gpurun_out/k3_roll_dram.csv python scripts/kernel_cases.py --case k3 --iters 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/k3_roll_dram.csv | cut -d, -f5,13- | tail -7
