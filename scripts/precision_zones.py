#!/usr/bin/env python
"""CPU study: which layers' bf16 OUTPUT rounding dominates the bf16 error (zone attribution)."""
import os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn
import oracle.model as om
from oracle.state import ModelConfig, make_state_dict

q = lambda t: t.bfloat16().float()
ZONE = ["top"]
ACTIVE = set()
ORIG = dict(conv3d=F.conv3d, linear=F.linear, conv_transpose3d=F.conv_transpose3d)
def wrap(fn):
    def f(x, w, b=None, *a, **k):
        y = fn(x, w, b, *a, **k)
        return q(y) if ZONE[-1] in ACTIVE else y
    return f
F.conv3d, F.linear, F.conv_transpose3d = wrap(ORIG["conv3d"]), wrap(ORIG["linear"]), wrap(ORIG["conv_transpose3d"])
def zoned(name, fn, namer=None):
    def g(*a, **k):
        ZONE.append(namer(*a, **k) if namer else name)
        try: return fn(*a, **k)
        finally: ZONE.pop()
    return g
om.window_attention = zoned("attn", om.window_attention)
om.ccf_ffn = zoned("ffn", om.ccf_ffn)
om.patch_merging = zoned("merge", om.patch_merging)
om.res_block = zoned("res", om.res_block, lambda sd, p, x: "res:" + p.split(".")[0])
om.channel_calibration = zoned("calib", om.channel_calibration)
om.projection_upsample = zoned("projup", om.projection_upsample)
_idwt = om.idwt_block
def idwt_block(sd, p, inp, skip, hf):
    ZONE.append("lf:" + p)
    low = F.conv3d(inp, sd[f"{p}.conv_lf_block.conv.weight"], padding=1)
    ZONE.pop()
    from oracle import haar
    rec = haar.waverec3((low,) + tuple(hf), "db1")
    return om.res_block(sd, f"{p}.conv_block", torch.cat((rec, skip), 1))
om.idwt_block = idwt_block

cfg = ModelConfig(img_size=(128,) * 3)
sd = make_state_dict(cfg, seed=0)
x = seeded_randn((1, 4, 128, 128, 128), 1)
torch.set_grad_enabled(False)
ref = om.waveformer_forward(sd, x, cfg)
def run(tag, zones):
    ACTIVE.clear(); ACTIVE.update(zones)
    y = om.waveformer_forward(sd, x, cfg); e = y - ref
    print(f"{tag:40s} max-rel {float(e.abs().max()/ref.abs().max()):.4f} rel-L2 {float(e.norm()/ref.norm()):.4f} argmax {float((y.argmax(1)==ref.argmax(1)).float().mean()):.5f}", flush=True)
allz = ["top","attn","ffn","merge","res:encoder1","res:encoder2","res:encoder3","res:encoder4","calib","lf:decoder4","lf:decoder3","lf:decoder2","res:decoder4","res:decoder3","res:decoder2","res:decoder1","projup"]
run("all", allz)
for z in allz: run(z, [z])

# ---- policy candidates: round GEMM inputs AND outputs inside the chosen zones ----
def wrap2(fn):
    def f(x, w, b=None, *a, **k):
        on = ZONE[-1] in ACTIVE
        if on: x, w = q(x), q(w)
        y = fn(x, w, b, *a, **k)
        return q(y) if on else y
    return f
F.conv3d, F.linear, F.conv_transpose3d = wrap2(ORIG["conv3d"]), wrap2(ORIG["linear"]), wrap2(ORIG["conv_transpose3d"])
convz = [z for z in allz if z.startswith(("res:", "lf:")) or z in ("calib", "projup")]
run("POLICY conv blocks bf16 in+out", convz)
run("POLICY conv blocks + top bf16", convz + ["top"])
run("POLICY conv blocks + ffn + merge bf16", convz + ["ffn", "merge"])
run("POLICY everything but attn", [z for z in allz if z != "attn"])
run("POLICY everything", allz)

# split the "top" zone by layer
def wrap3(fn, kind):
    def f(x, w, b=None, *a, **k):
        z = ZONE[-1]
        if z == "top":
            z = {"conv3d": "top:patch_embed" if w.shape[-1] == 2 else "top:out", "conv_transpose3d": "top:transp"}.get(kind, z)
        on = z in ACTIVE
        if on: x, w = q(x), q(w)
        y = fn(x, w, b, *a, **k)
        return q(y) if on else y
    return f
F.conv3d, F.linear, F.conv_transpose3d = wrap3(ORIG["conv3d"], "conv3d"), wrap3(ORIG["linear"], "linear"), wrap3(ORIG["conv_transpose3d"], "conv_transpose3d")
base = convz + ["ffn", "merge"]
run("POLICY2 base + transp + out", base + ["top:transp", "top:out"])
run("POLICY2 base + patch_embed", base + ["top:patch_embed"])
run("POLICY2 only patch_embed", ["top:patch_embed"])
