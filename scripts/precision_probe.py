#!/usr/bin/env python
"""bf16 accuracy of the product forward against the fp32 reference fixture (tests/golden/waveformer_128.npz):
max-relative error, relative L2 error and argmax agreement, plus where the error enters (per-stage)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_npz, seeded_randn  # noqa: E402
from oracle.model import waveformer_forward  # noqa: E402
from oracle.state import ModelConfig, make_state_dict  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = ModelConfig(img_size=(128,) * 3)
sd = make_state_dict(cfg, seed=0)
x = seeded_randn((1, 4, 128, 128, 128), 1)
with torch.no_grad():
    ref, mid = waveformer_forward(sd, x, cfg, return_intermediates=True)


def report(tag, y):
    y = y.float().cpu()
    err = (y - ref)
    print(f"{tag:34s} max-rel {float(err.abs().max() / ref.abs().max()):.4f}  rel-L2 {float(err.norm() / ref.norm()):.4f}  "
          f"argmax-agree {float((y.argmax(1) == ref.argmax(1)).float().mean()):.5f}")


def build(dtype):
    m = Waveformer(**cfg.kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    return m.cuda().to(dtype).to(memory_format=torch.channels_last_3d)


with torch.no_grad():
    m32 = build(torch.float32)
    report("fp32 product", m32(x.cuda()))
    m16 = build(torch.bfloat16)
    report("bf16 product (pure bf16)", m16(x.cuda()))
    # where does it enter?  feed fp32 encoder outputs into the bf16 decoder and vice versa
    outs32, hf32 = m32.waveformer_encoder(x.cuda().contiguous(memory_format=torch.channels_last_3d))
    outs16, hf16 = m16.waveformer_encoder(x.cuda().bfloat16().contiguous(memory_format=torch.channels_last_3d))
    for i, (a, b) in enumerate(zip(outs16, outs32)):
        e = (a.float() - b)
        print(f"  encoder out{i}: max-rel {float(e.abs().max() / b.abs().max()):.4f} rel-L2 {float(e.norm() / b.norm()):.4f}")
    # autocast-style: fp32 weights + autocast(bf16)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        try:
            report("fp32 model under autocast(bf16)", m32(x.cuda()))
        except Exception as exc:  # noqa: BLE001
            print("autocast run failed:", repr(exc)[:300])
