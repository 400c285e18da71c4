#!/usr/bin/env python
"""CUDA-event time per sub-module of one Waveformer forward (16-bit policy; --batch N windows, default 2): where the step goes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import prepare_inference  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

torch.manual_seed(0)
if os.environ.get("WF_CUDNN_BENCHMARK"):
    torch.backends.cudnn.benchmark = True
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4,
                                 feat_size=[48, 96, 192, 384], num_heads=[3, 6, 12, 24]).eval().cuda(), torch.bfloat16)
BATCH = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 2
x = torch.randn(BATCH, 4, 128, 128, 128, device="cuda").contiguous(memory_format=torch.channels_last_3d)
names = {}
for n, mod in m.named_modules():
    depth = n.count(".")
    if n and (depth == 0 or (n.startswith("waveformer_encoder.") and depth <= 2) or
              (depth == 1 and not n.startswith("waveformer_encoder")) or n.endswith(".mlp") or n.endswith(".attn")):
        names[mod] = n
rec = {}


def pre(mod, inp):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    rec.setdefault(names[mod], []).append([e, None])


def post(mod, inp, out):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    rec[names[mod]][-1][1] = e


for mod in names:
    mod.register_forward_pre_hook(pre)
    mod.register_forward_hook(post)
with torch.no_grad():
    for _ in range(3):
        m(x)
    rec.clear()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        m(x)
    b.record()
    torch.cuda.synchronize()
print(f"forward (with hooks): {a.elapsed_time(b) / 5:.2f} ms")
for n, evs in rec.items():
    ms = sum(s.elapsed_time(e) for s, e in evs) / 5
    if ms > 0.05:
        print(f"{ms:8.3f} ms  {n}")
