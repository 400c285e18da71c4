#!/usr/bin/env python
"""Which source lines of the package launch torch's own elementwise kernels (copies, casts, adds) in one forward?
Torch profiler with python stacks; prints device time of aten ops grouped by the innermost waveformer_b200 frame."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import prepare_inference  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

torch.manual_seed(0)
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4,
                                 feat_size=[48, 96, 192, 384], num_heads=[3, 6, 12, 24], drop_path_rate=0.1).eval().cuda(),
                      torch.bfloat16)
x = torch.randn(2, 4, 128, 128, 128, device="cuda").contiguous(memory_format=torch.channels_last_3d)
with torch.no_grad():
    for _ in range(3):
        m(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        m(x)
        torch.cuda.synchronize()
want = ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::cat", "aten::fill_", "aten::zero_", "aten::sub",
        "aten::div", "aten::gelu", "aten::native_layer_norm", "aten::index", "aten::clone")
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.name not in want:
        continue
    t = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    if t <= 0:
        continue
    where = "?"
    for fr in ev.stack or []:
        if "waveformer_b200" in fr and "scripts" not in fr:
            where = fr.split("waveformer_b200/")[-1]
            break
    agg[(ev.name, where)][0] += 1
    agg[(ev.name, where)][1] += t
tot = sum(v[1] for v in agg.values())
print(f"torch elementwise ops in one forward: {tot:.0f} us")
for (name, where), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:50]:
    print(f"{t:8.1f} us x{c:3d} {name:14s} {where}")
