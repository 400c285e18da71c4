#!/bin/bash
# time conv3d_k3_c48 for every prebuilt parameter variant under waveformer_b200/lib/variants (ring / lag / staging rows)
mkdir -p gpurun_out
cp waveformer_b200/lib/libwaveformer_b200.so /tmp/lib_main.so
for v in waveformer_b200/lib/variants/*.so; do
  cp $v waveformer_b200/lib/libwaveformer_b200.so
  echo "== $(basename $v)"
  timeout 120 python scripts/kernel_cases.py --case k3 --iters 10 2>&1 | grep conv3d_k3
done
cp /tmp/lib_main.so waveformer_b200/lib/libwaveformer_b200.so
