#!/bin/bash
# round 2, pass B: GPU test suite + precision report
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
timeout 600 python scripts/precision_probe.py storage > gpurun_out/precision.log 2>&1
tail -25 gpurun_out/pytest.log
cat gpurun_out/precision.log | tail -20
