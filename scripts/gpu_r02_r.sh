#!/bin/bash
# round-2 closing evidence: full GPU test suite, the bench line, warmed launch lists at 6 and 2 windows per forward
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
tail -3 gpurun_out/pytest.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
for b in 6 2; do
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward_b$b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b$b.csv \
    python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward_b$b.log 2>&1
tail -1 gpurun_out/plain_forward_b$b.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
