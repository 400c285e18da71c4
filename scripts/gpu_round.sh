#!/bin/bash
# One GPU-box pass: tests, smoke, bench, then profiler captures (each ncu run follows a plain run of the same command).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
(timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --rows 70 > gpurun_out/prof_bf16.log 2>&1
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_forward_bf16.csv \
    python scripts/profile_forward.py --dtype bf16 --batch 2 --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward.log 2>&1
timeout 300 python scripts/haar_config2.py --dtype bf16 > gpurun_out/haar_bf16.log 2>&1
timeout 300 python scripts/haar_config2.py --dtype bf16 --iters 2 > gpurun_out/plain_haar.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dwt|idwt' -s 4 -c 4 -f -o gpurun_out/haar_bf16 \
    python scripts/haar_config2.py --dtype bf16 --iters 2 > gpurun_out/ncu_haar.log 2>&1
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; tail -c 600 gpurun_out/bench.log; head -3 gpurun_out/prof_bf16.log; cat gpurun_out/haar_bf16.log
