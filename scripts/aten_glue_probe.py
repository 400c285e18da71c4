#!/usr/bin/env python
"""Which torch (ATen) elementwise launches are left in the window forward, by op, input shapes and Python call site."""
import os, sys, collections
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import prepare_inference
from waveformer_b200.network_models import Waveformer
torch.manual_seed(0)
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4, feat_size=[48, 96, 192, 384],
                                 num_heads=[3, 6, 12, 24], drop_path_rate=0.1).eval().cuda(), torch.bfloat16)
x = torch.randn(2, 4, 128, 128, 128, device="cuda").contiguous(memory_format=torch.channels_last_3d)
with torch.no_grad():
    for _ in range(2):
        m(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        m(x)
        torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True, group_by_stack_n=6):
    if e.key.startswith("aten::") and e.device_time_total > 0 and e.key in ("aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::relu",
            "aten::mean", "aten::sigmoid", "aten::_to_copy", "aten::contiguous", "aten::clone", "aten::cat", "aten::gelu", "aten::fill_", "aten::zero_"):
        stack = [s for s in e.stack if "waveformer_b200" in s or "site-packages/torch/nn/modules" in s][:3]
        rows.append((e.device_time_total, e.count, e.key, str(e.input_shapes)[:90], " <- ".join(s.split("/")[-1][:60] for s in stack)))
rows.sort(reverse=True)
for t, c, k, sh, st in rows[:45]:
    print(f"{t:9.1f} us x{c:3d} {k:16s} {sh:90s} {st}")
