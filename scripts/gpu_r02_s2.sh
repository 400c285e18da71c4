#!/bin/bash
mkdir -p gpurun_out
WF_UP1_RING=1 timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn | sed 's/^/RING1 /' | tee -a gpurun_out/upsample_probe.log
timeout 300 python scripts/upsample_probe.py 2>&1 | grep -v Warn | tee -a gpurun_out/upsample_probe.log
