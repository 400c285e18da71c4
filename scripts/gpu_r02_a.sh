#!/bin/bash
# round 2, pass A: GPU test suite, smoke, precision report, module timings, short bench.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
timeout 600 python scripts/precision_probe.py > gpurun_out/precision.log 2>&1
timeout 300 python scripts/module_times.py > gpurun_out/module_times.log 2>&1
(timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
tail -15 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log
cat gpurun_out/precision.log | tail -20
grep -E "forward|block1.0|encoder[1234]$|learnable_up[34]$|decoder1$|waveformer_encoder$" gpurun_out/module_times.log
tail -c 1500 gpurun_out/bench.log
