#!/bin/bash
# One B200 pass: GPU test suite, smoke, per-kernel timings, per-module timings, precision report, bench.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_check.sh'
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
for c in haar c4 k3 dwconv convt head layernorm attn; do timeout 300 python scripts/kernel_cases.py --case $c --iters 10; done > gpurun_out/kernel_cases.log 2>&1
timeout 300 python scripts/module_times.py > gpurun_out/module_times.log 2>&1
timeout 600 python scripts/precision_probe.py > gpurun_out/precision.log 2>&1
(timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/bench.log)
tail -4 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/kernel_cases.log
grep -E "forward|block1.0|encoder[1234]$|learnable_up[34]$|decoder1$|waveformer_encoder$" gpurun_out/module_times.log
grep -E "policy, attention fp16|weights" gpurun_out/precision.log; tail -c 300 gpurun_out/bench.log
