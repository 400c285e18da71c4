#!/usr/bin/env python
"""Time the InstanceNorm passes of the 128^3 residual blocks at the bench's batching (6 windows, fp16, 1.2 GB per tensor),
and the first block's convolution with / without the stored shortcut."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import ops  # noqa: E402

g = torch.Generator("cuda").manual_seed(0)
B = int(os.environ.get("PROBE_B", "6"))
x = torch.randn(B, 128, 128, 128, 48, device="cuda", generator=g).half().permute(0, 4, 1, 2, 3)
r = torch.randn(B, 128, 128, 128, 48, device="cuda", generator=g).half().permute(0, 4, 1, 2, 3)
st, rst = ops.instance_norm_stats(x), ops.instance_norm_stats(r)
fns = {"stats": lambda: ops.instance_norm_stats(x),
       "apply": lambda: ops.instance_norm_act(x, "leakyrelu", 0.01, stats=st),
       "apply+res": lambda: ops.instance_norm_act(x, "leakyrelu", 0.01, res=r, res_norm=True, stats=st, res_stats=rst)}
xin = torch.randn(B, 128, 128, 128, 4, device="cuda", generator=g).half().permute(0, 4, 1, 2, 3)
w1 = (torch.randn(48, 4, 3, 3, 3, device="cuda", generator=g) * 0.1).half()
w3 = (torch.randn(48, 4, 1, 1, 1, device="cuda", generator=g) * 0.5).half()
fns["apply+shortcut4"] = lambda: ops.instance_norm_act_shortcut4(x, xin, w3, st, rst)
fns["conv_c4 (shortcut stored)"] = lambda: ops.conv3d_c4_in_stats(xin, w1, w3)
fns["conv_c4 (shortcut statistics only)"] = lambda: ops.conv3d_c4_in_stats(xin, w1, w3, store_shortcut=False)
tag = f"B={B}"
for name, fn in fns.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    print(f"{tag} {name}: median {ts[5] * 1000:.1f} us  min {ts[0] * 1000:.1f} us")
