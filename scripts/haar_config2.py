#!/usr/bin/env python
"""BASELINE configs[1]: isolated DWT3D -> IDWT3D round trip on a 2x48x128^3 feature map (the ncu target).

    python scripts/haar_config2.py [--dtype bf16|f32] [--layout ncdhw|ndhwc] [--iters 5]

Prints CUDA-event times and algorithmic GB/s per kernel; run under `ncu --set full -k regex:dwt|idwt` for the DRAM
counters (the numbers printed under ncu are not bench values).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--layout", default="ncdhw")
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
x = torch.randn((2, 48, 128, 128, 128), device="cuda", generator=torch.Generator("cuda").manual_seed(0)).to(dtype)
if args.layout == "ndhwc":
    x = x.view(2, 128, 128, 128, 48)
    dwt, idwt = ops._dwt_ndhwc_raw, ops._idwt_ndhwc_raw
else:
    dwt, idwt = ops._dwt_ncdhw_raw, ops._idwt_ncdhw_raw
alg = 2 * x.numel() * x.element_size()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.iters


ll, hf = dwt(x, True)
t1 = timed(lambda: dwt(x, True))
t2 = timed(lambda: idwt(ll, hf))
y = idwt(ll, hf)
err = float((y.float() - x.float()).abs().max())
print(f"{args.layout} {args.dtype}: dwt {t1:.4f} ms {alg / t1 / 1e6:.0f} GB/s | idwt {t2:.4f} ms {alg / t2 / 1e6:.0f} GB/s | "
      f"round-trip max-abs err {err:.3e} | algorithmic bytes per launch {alg}")
