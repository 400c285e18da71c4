#!/usr/bin/env python
"""CPU study of where bf16 rounding hurts: the oracle forward with bf16 rounding injected at chosen points.
Not part of the product or the tests; used to choose the mixed-precision policy documented in DESIGN.md."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn  # noqa: E402
import oracle.model as om  # noqa: E402
from oracle import haar  # noqa: E402
from oracle.state import ModelConfig, make_state_dict  # noqa: E402

q = lambda t: t.bfloat16().float()
ORIG = dict(conv3d=F.conv3d, linear=F.linear, conv_transpose3d=F.conv_transpose3d, layer_norm=F.layer_norm,
            instance_norm=F.instance_norm, group_norm=F.group_norm, gelu=F.gelu, leaky_relu=F.leaky_relu,
            interpolate=F.interpolate, wavedec3=haar.wavedec3, waverec3=haar.waverec3, softmax=torch.softmax)


def install(gemm_in=False, gemm_out=False, norm_out=False, act_out=False, interp_out=False, dwt_in=False, dwt_out=False,
            resid=False):
    def wrap_gemm(fn):
        def f(x, w, b=None, *a, **k):
            if gemm_in:
                x, w = q(x), q(w)
            y = fn(x, w, b, *a, **k)
            return q(y) if gemm_out else y
        return f
    F.conv3d = wrap_gemm(ORIG["conv3d"])
    F.linear = wrap_gemm(ORIG["linear"])
    F.conv_transpose3d = wrap_gemm(ORIG["conv_transpose3d"])
    wrap_out = lambda fn, on: (lambda *a, **k: q(fn(*a, **k))) if on else fn
    F.layer_norm = wrap_out(ORIG["layer_norm"], norm_out)
    F.instance_norm = wrap_out(ORIG["instance_norm"], norm_out)
    F.group_norm = wrap_out(ORIG["group_norm"], norm_out)
    F.gelu = wrap_out(ORIG["gelu"], act_out)
    F.leaky_relu = wrap_out(ORIG["leaky_relu"], act_out)
    F.interpolate = wrap_out(ORIG["interpolate"], interp_out)

    def wd(x, *a, **k):
        c = ORIG["wavedec3"](q(x) if dwt_in else x, *a, **k)
        if dwt_out:
            c = (q(c[0]),) + tuple({kk: q(v) for kk, v in d.items()} for d in c[1:])
        return c
    haar.wavedec3 = wd
    om.RESID_Q = resid


cfg = ModelConfig(img_size=(128,) * 3)
sd = make_state_dict(cfg, seed=0)
x = seeded_randn((1, 4, 128, 128, 128), 1)
torch.set_grad_enabled(False)
install()
ref = om.waveformer_forward(sd, x, cfg)


def run(tag, **kw):
    install(**kw)
    t = time.time()
    y = om.waveformer_forward(sd, x, cfg)
    e = y - ref
    print(f"{tag:60s} max-rel {float(e.abs().max() / ref.abs().max()):.4f} rel-L2 {float(e.norm() / ref.norm()):.4f} "
          f"argmax {float((y.argmax(1) == ref.argmax(1)).float().mean()):.5f}  ({time.time() - t:.0f}s)", flush=True)


run("A  gemm inputs bf16 (floor of bf16 tensor-core math)", gemm_in=True)
run("B  A + gemm outputs bf16", gemm_in=True, gemm_out=True)
run("C  B + norm outputs bf16", gemm_in=True, gemm_out=True, norm_out=True)
run("D  C + act/interp outputs + dwt in/out bf16 (~pure bf16)", gemm_in=True, gemm_out=True, norm_out=True, act_out=True,
    interp_out=True, dwt_in=True, dwt_out=True)
run("E  only dwt input bf16 (LN out rounded before DWT)", dwt_in=True)
run("F  only dwt outputs bf16", dwt_out=True)
run("G  only norm outputs bf16", norm_out=True)
run("H  only gemm outputs bf16", gemm_out=True)
