#!/usr/bin/env python
"""CPU study: which rounding point inside the bf16 window attention drives the logit error (emulates attn_tc.cu)."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn
import oracle.model as om
from oracle.state import ModelConfig, make_state_dict

q = lambda t: t.bfloat16().float()
K = dict(x=False, w=False, qkv=False, p=False, o=False, out=False, stages=(1, 2, 3, 4))

def emu(sd, p, x, heads):
    stage = int(p.split("block")[1][0])
    on = lambda k: K[k] and stage in K["stages"]
    b_, n, c = x.shape
    hd = c // heads
    wq, bq, wp, bp = sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"], sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"]
    if on("x"): x = q(x)
    if on("w"): wq, bq, wp, bp = q(wq), q(bq), q(wp), q(bp)
    qkv = F.linear(x, wq, bq).reshape(b_, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    qq, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    if on("qkv"): qq, k, v = q(qq), q(k), q(v)
    s = qq @ k.transpose(-2, -1) + om.relative_position_bias(sd[f"{p}.relative_position_bias_table"], sd[f"{p}.relative_position_index"])[None]
    s = s - s.amax(-1, keepdim=True)
    e = torch.exp(s)
    den = e.sum(-1, keepdim=True)
    if on("p"): e = q(e)
    o = ((e @ v) / den).transpose(1, 2).reshape(b_, n, c)
    if on("o"): o = q(o)
    y = F.linear(o, wp, bp)
    return q(y) if on("out") else y

om.window_attention = emu
cfg = ModelConfig(img_size=(128,) * 3)
sd = make_state_dict(cfg, seed=0)
x = seeded_randn((1, 4, 128, 128, 128), 1)
torch.set_grad_enabled(False)
ref = om.waveformer_forward(sd, x, cfg)
def run(tag, **kw):
    K.update(dict(x=False, w=False, qkv=False, p=False, o=False, out=False, stages=(1, 2, 3, 4))); K.update(kw)
    y = om.waveformer_forward(sd, x, cfg); e = y - ref
    print(f"{tag:36s} max-rel {float(e.abs().max()/ref.abs().max()):.4f} rel-L2 {float(e.norm()/ref.norm()):.4f} argmax {float((y.argmax(1)==ref.argmax(1)).float().mean()):.5f}", flush=True)
A = dict(x=True, w=True, qkv=True, p=True, o=True, out=True)
run("fp32 (sanity)")
run("all points", **A)
for k in A: run(f"only {k}", **{k: True})
run("all but out", **{**A, "out": False})
run("all but out,x", **{**A, "out": False, "x": False})
for s in (1, 2, 3, 4): run(f"all points, stage {s} only", **A, stages=(s,))
