#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log)
timeout 600 python scripts/precision_probe.py > gpurun_out/precision.log 2>&1
timeout 300 python scripts/haar_config2.py --dtype bf16 > gpurun_out/haar_bf16.log 2>&1
timeout 300 python scripts/haar_config2.py --dtype f32 >> gpurun_out/haar_bf16.log 2>&1
timeout 300 python scripts/haar_config2.py --dtype bf16 --layout ndhwc >> gpurun_out/haar_bf16.log 2>&1
timeout 300 python scripts/profile_forward.py --dtype bf16 --batch 2 --rows 70 > gpurun_out/prof_bf16.log 2>&1
(timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log)
tail -15 gpurun_out/pytest.log; cat gpurun_out/precision.log; cat gpurun_out/haar_bf16.log; tail -2 gpurun_out/smoke.log; grep "forward batch" gpurun_out/prof_bf16.log
