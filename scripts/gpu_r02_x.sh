#!/bin/bash
# closing run of the final build: bench line (all extras), smoke, refreshed launch list at six windows per forward
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
for b in 6 2; do
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/plain_forward_b$b.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_forward_b$b.csv \
    python scripts/profile_forward.py --dtype bf16 --batch $b --no-profiler --iters 1 --warm 1 > gpurun_out/ncu_forward_b$b.log 2>&1
tail -1 gpurun_out/plain_forward_b$b.log
done
