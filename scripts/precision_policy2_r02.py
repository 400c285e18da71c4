#!/usr/bin/env python
"""CPU study (round 2, part 2): what is left after the all-fp16 storage policy - the attention operands and the skip
blocks encoder2..4 - and which cheap counter-measure removes it (split-fp16 S path; fp32 shortcut / fp32 conv results)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn
import oracle.model as om
from oracle.state import ModelConfig, make_state_dict

h = lambda t: t.half().float()
ORIG = dict(conv3d=F.conv3d, linear=F.linear)
ZONE = ["top"]
MODE = {}


def conv3d(x, w, b=None, *a, **k):
    z = ZONE[-1]
    m = MODE.get(z)
    if m is None:
        return ORIG["conv3d"](x, w, b, *a, **k)
    if m == "ops":          # operands rounded, fp32 result
        return ORIG["conv3d"](h(x), h(w), b, *a, **k)
    if m == "out":          # exact operands, result rounded
        return h(ORIG["conv3d"](x, w, b, *a, **k))
    return h(ORIG["conv3d"](h(x), h(w), b, *a, **k))


F.conv3d = conv3d


def zoned(name, fn, namer=None):
    def g(*a, **k):
        ZONE.append(namer(*a, **k) if namer else name)
        try:
            return fn(*a, **k)
        finally:
            ZONE.pop()
    return g


_res = om.res_block


def res_block(sd, p, x):
    name = "res:" + p.split(".")[0]
    m = MODE.get(name + ":in")
    if m == "fp16":
        x = h(x)
    ZONE.append(name)
    try:
        return _res(sd, p, x)
    finally:
        ZONE.pop()


om.res_block = res_block
ATTN = {"mode": "fp32"}


def window_attention(sd, p, x, heads):
    b_, n, c = x.shape
    hd = c // heads
    scale = hd ** -0.5
    mode = ATTN["mode"]
    W, B = sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"]
    if mode == "fp16":                       # today: fp16 operands everywhere
        qkv = ORIG["linear"](h(x), h(W), B)
    else:                                    # split-fp16 QKV GEMM (hi/lo operands, 3 MMAs): ~fp32
        qkv = ORIG["linear"](x, W, B)
    qkv = qkv.reshape(b_, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    if mode == "fp16":
        q, k = h(q), h(k)
    elif mode == "split":                    # q = hi + lo, k = hi + lo; S = qh kh + ql kh + qh kl (drops ql kl)
        qh, kh = h(q), h(k)
        ql, kl = h(q - qh), h(k - kh)
        s = qh @ kh.transpose(-2, -1) + ql @ kh.transpose(-2, -1) + qh @ kl.transpose(-2, -1)
    if mode != "split":
        s = q @ k.transpose(-2, -1)
    s = s + om.relative_position_bias(sd[f"{p}.relative_position_bias_table"], sd[f"{p}.relative_position_index"])[None]
    a = torch.softmax(s, dim=-1)
    if mode != "fp32":
        a, v = h(a), h(v)
    o = (a @ v).transpose(1, 2).reshape(b_, n, c)
    if mode != "fp32":
        return ORIG["linear"](h(o), h(sd[f"{p}.proj.weight"]), sd[f"{p}.proj.bias"])
    return ORIG["linear"](o, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


om.window_attention = window_attention

cfg = ModelConfig(img_size=(128,) * 3)
sd = make_state_dict(cfg, seed=0)
x = seeded_randn((1, 4, 128, 128, 128), 1)
torch.set_grad_enabled(False)
ref = om.waveformer_forward(sd, x, cfg)


def run(tag, mode, attn="fp32"):
    MODE.clear(); MODE.update(mode); ATTN["mode"] = attn
    y = om.waveformer_forward(sd, x, cfg); e = y - ref
    same = (y.argmax(1) == ref.argmax(1)).float().mean()
    print(f"{tag:58s} max-rel {float(e.abs().max()/ref.abs().max()):.5f} rel-L2 {float(e.norm()/ref.norm()):.5f} argmax {float(same):.5f}", flush=True)


run("attention fp16 operands (q, k, v, P, o; today)", {}, "fp16")
run("attention split-fp16 S path, fp16 P / v / o", {}, "split")
for blk in ("res:encoder2", "res:encoder3"):
    run(f"{blk}: input rounded only", {blk + ":in": "fp16"})
    run(f"{blk}: conv operands rounded, fp32 results", {blk: "ops"})
    run(f"{blk}: conv results rounded only", {blk: "out"})
    run(f"{blk}: all", {blk: "all", blk + ":in": "fp16"})
