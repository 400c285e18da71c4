#!/usr/bin/env python
"""bf16 accuracy of the product forward on one 1x4x128^3 window, per precision policy and per weight set:
max-relative error, relative L2 error, argmax agreement, and argmax agreement restricted to voxels whose fp32 decision
margin (top-1 minus top-2 logit) exceeds the 2e-2 * max|logit| error tolerance.

Weight sets: "unit-gain" = oracle.state.make_state_dict (every layer ~unit gain: the stress case the parity tests use;
reference = CPU oracle); "ctor-init" = the constructor's own initialisation (what the reference calls random init:
trunc-normal 0.02 linears; reference = the fp32 product path, itself <= 1e-6 from the oracle)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn  # noqa: E402
from oracle.model import waveformer_forward  # noqa: E402
from oracle.state import ModelConfig, make_state_dict  # noqa: E402
from waveformer_b200 import prepare_inference  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
cfg = ModelConfig(img_size=(128,) * 3)
x = seeded_randn((1, 4, 128, 128, 128), 1)


def report(tag, y, ref):
    y = y.float().cpu()
    err = y - ref
    tol = 2e-2 * float(ref.abs().max())
    top = ref.topk(2, dim=1).values
    clear = (top[:, 0] - top[:, 1]) > tol
    same = y.argmax(1) == ref.argmax(1)
    print(f"{tag:46s} max-rel {float(err.abs().max() / ref.abs().max()):.4f}  rel-L2 {float(err.norm() / ref.norm()):.4f}  "
          f"argmax {float(same.float().mean()):.5f}  argmax|margin>tol {float(same[clear].float().mean()):.5f} "
          f"({float(clear.float().mean()):.3f} of voxels)", flush=True)


def build(sd, **policy):
    m = Waveformer(**cfg.kwargs()).eval()
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    return m


POLICIES = [("pure bf16 (model.to(bf16))", dict(fp32_stream=False)),
            ("r01 policy: bf16 storage, attn fp16", dict(storage="bf16", attention="fp16")),
            ("fp16 storage, attn fp16", dict(attention="fp16")),
            ("r02 policy: fp16 storage, attn fp16x2 (default)", dict()),
            ("r02 policy without the skip blocks' input pair", dict(skip_split_input=False)),
            ("fp16 storage, attention fp32 CUDA cores", dict(attention="fp32")),
            ("bf16 storage, attention fp32 CUDA cores", dict(storage="bf16", attention="fp32"))]
if len(sys.argv) > 1:
    POLICIES = [p for p in POLICIES if any(a in p[0] for a in sys.argv[1:])]

with torch.no_grad():
    for name in ("unit-gain", "ctor-init"):
        if name == "unit-gain":
            sd = make_state_dict(cfg, seed=0)
            ref = waveformer_forward(sd, x, cfg)
        else:
            torch.manual_seed(0)
            sd = {k: v.clone() for k, v in Waveformer(**cfg.kwargs()).state_dict().items()}
            ref = None
        m32 = prepare_inference(build(sd).cuda(), torch.float32)
        y32 = m32(x.cuda()).float().cpu()
        if ref is None:
            ref = y32
        print(f"== weights: {name}")
        report("fp32 product", y32, ref)
        del m32
        for tag, kw in POLICIES:
            m = prepare_inference(build(sd).cuda(), torch.bfloat16, **kw)
            report(tag, m(x.cuda()), ref)
            del m
            torch.cuda.empty_cache()
