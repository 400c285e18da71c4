#!/usr/bin/env python
"""Time the fused multi-level trilinear upsampling + shortcut of the stage-1 / stage-2 blocks and ProjectionUpsample's nn.Upsample
at six windows per forward (WF_AB_OLD=1: the cell kernel; default: the z-walking kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import ops  # noqa: E402

g = torch.Generator("cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
B = 6
s1 = [rn(B, n, n, n, 48).half() for n in (32, 16, 8)]
b1 = rn(B, 64, 64, 64, 48)
s2 = [rn(B, n, n, n, 96).half() for n in (16, 8)]
b2 = rn(B, 32, 32, 32, 96)
u4 = rn(B, 32, 32, 32, 192).half()
u3 = rn(B, 32, 32, 32, 96).half()
fns = {"stage 1: 3 sources -> 64^3 x 48 fp32 + base": lambda: ops.upsample_trilinear_add(s1, (64, 64, 64), base=b1, out_dtype=torch.float32),
       "stage 2: 2 sources -> 32^3 x 96 fp32 + base": lambda: ops.upsample_trilinear_add(s2, (32, 32, 32), base=b2, out_dtype=torch.float32),
       "learnable_up4: 32^3 -> 64^3 x 192 fp16, align_corners": lambda: ops.upsample_trilinear_add([u4], (64, 64, 64), align_corners=True),
       "learnable_up3: 32^3 -> 64^3 x 96 fp16, align_corners": lambda: ops.upsample_trilinear_add([u3], (64, 64, 64), align_corners=True)}
ref = {}
for name, fn in fns.items():
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    print(f"AB_OLD={os.environ.get('WF_AB_OLD', '0')} {name}: median {ts[5] * 1000:.1f} us  checksum {float(out.float().double().sum()):.6f} {float(out.float().abs().double().sum()):.6f}")
