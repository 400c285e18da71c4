#!/usr/bin/env python
"""Window attention on the four stage geometries of one batch-2 window forward: CUDA-event time per call, clocks per
128x512 score tile against the 4096-clk MUFU floor (65536 ex2 at 16/clk/SM), achieved TFLOP/s.

    python scripts/attn_bench.py [--fmt fp16|bf16] [--iters 20] [--only LABEL]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200.network_models import Attention  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fmt", default="fp16")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--only", default="")
args = ap.parse_args()
fmt = torch.float16 if args.fmt == "fp16" else torch.bfloat16
SM, MHZ = 148, 1965.0
# (label, batch, grid edge, C, heads): every attention call of one forward at sw_batch 2 (x2 blocks per stage)
CASES = [("s1.L1", 2, 32, 48, 3), ("s1.L2", 2, 16, 48, 3), ("s1.L3", 2, 8, 48, 3), ("s2.L1", 2, 16, 96, 6),
         ("s2.L2", 2, 8, 96, 6), ("s3.L1", 2, 8, 192, 12), ("s4", 2, 8, 384, 24)]
tot = 0.0
for label, b, g, c, h in CASES:
    if args.only and args.only != label:
        continue
    torch.manual_seed(0)
    att = Attention(c, num_heads=h, qkv_bias=True, window_size=8).cuda().eval()
    att.compute_dtype, att.out_dtype = fmt, torch.float32
    x = torch.randn(b, g, g, g, c, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            att.forward_grid(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            att.forward_grid(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    windows = b * (g // 8) ** 3
    tiles = windows * h * 4
    waves = -(-tiles // min(SM, max(1, (SM // (h * 4)) * h * 4)))
    flops = windows * (4096 * c * c + 1048576 * c)
    clk_tile = ms * 1e-3 * MHZ * 1e6 / max(1, -(-tiles // SM))
    tot += ms
    print(f"{label:6s} B={b} grid={g}^3 C={c:3d} heads={h:2d} windows={windows:4d} tiles={tiles:5d}: {ms * 1e3:8.1f} us/call "
          f"(3 launches)  {flops / ms / 1e9:7.1f} TFLOP/s  ~{clk_tile:7.0f} clk per tile-wave (MUFU floor 4096)")
print(f"sum over the 7 geometries: {tot * 1e3:.1f} us (x2 blocks per stage = one forward's attention)")
