#!/usr/bin/env python
"""Kernel-level time breakdown of one Waveformer forward (torch profiler / CUPTI; no nsys in the image).

    python scripts/profile_forward.py [--dtype bf16|f32] [--batch 2] [--out gpurun_out/profile_forward.txt]
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200.network_models import Waveformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--out", default="gpurun_out/profile_forward.txt")
ap.add_argument("--rows", type=int, default=45)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--skip-blocks", default="fp16")
ap.add_argument("--no-profiler", action="store_true", help="plain timed forwards only (the command ncu wraps)")
args = ap.parse_args()
dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
torch.manual_seed(0)
from waveformer_b200 import prepare_inference  # noqa: E402
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4,
                                 feat_size=[48, 96, 192, 384], num_heads=[3, 6, 12, 24], drop_path_rate=0.1).eval().cuda(), dtype,
                      **({"skip_blocks": args.skip_blocks} if dtype == torch.bfloat16 else {}))
x = torch.randn(args.batch, 4, 128, 128, 128, device="cuda").contiguous(memory_format=torch.channels_last_3d)  # fp32 window, as the inferer gathers it
with torch.no_grad():
    for _ in range(args.warm):
        m(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        m(x)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.iters
    if args.no_profiler:
        print(f"forward batch={args.batch} dtype={args.dtype}: {ms:.2f} ms")
        sys.exit(0)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        m(x)
        torch.cuda.synchronize()
table = prof.key_averages().table(sort_by="cuda_time_total", row_limit=args.rows, max_name_column_width=90)
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as f:
    f.write(f"forward batch={args.batch} dtype={args.dtype}: {ms:.2f} ms (CUDA events, 5 iters)\n")
    f.write(table)
print(f"forward batch={args.batch} dtype={args.dtype}: {ms:.2f} ms")
print("\n".join(table.splitlines()[:args.rows + 8]))
