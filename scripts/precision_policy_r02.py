#!/usr/bin/env python
"""CPU study (round 2): label agreement of candidate 16-bit storage policies, simulated on the oracle by rounding the
operands and results of every GEMM-like op (conv3d / linear / conv_transpose3d) of a zone to that zone's format.

    python scripts/precision_policy_r02.py [unit|init]
"""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn
import oracle.model as om
from oracle.state import ModelConfig, make_state_dict

Q = {"bf16": lambda t: t.bfloat16().float(), "fp16": lambda t: t.half().float(), "fp32": lambda t: t}
ZONE = ["top"]
POLICY = {}
ORIG = dict(conv3d=F.conv3d, linear=F.linear, conv_transpose3d=F.conv_transpose3d)


def wrap(fn, kind):
    def f(x, w, b=None, *a, **k):
        z = ZONE[-1]
        if z == "top":
            z = {"conv3d": "top:patch_embed" if w.shape[-1] == 2 else "top:out", "conv_transpose3d": "top:transp"}.get(kind, z)
        fmt = POLICY.get(z, "fp32")
        spec = fmt.split(">")          # "bf16>fp32" = bf16 operands, fp32 (unrounded) result
        qi, qo = Q[spec[0]], Q[spec[-1]]
        return qo(fn(qi(x), qi(w), b, *a, **k))
    return f


F.conv3d, F.linear, F.conv_transpose3d = (wrap(ORIG[k], k) for k in ("conv3d", "linear", "conv_transpose3d"))


def zoned(name, fn, namer=None):
    def g(*a, **k):
        ZONE.append(namer(*a, **k) if namer else name)
        try:
            return fn(*a, **k)
        finally:
            ZONE.pop()
    return g


om.window_attention = zoned("attn", om.window_attention)
om.ccf_ffn = zoned("ffn", om.ccf_ffn)
om.patch_merging = zoned("merge", om.patch_merging)
om.res_block = zoned("res", om.res_block, lambda sd, p, x: "res:" + p.split(".")[0])
om.channel_calibration = zoned("calib", om.channel_calibration)
om.projection_upsample = zoned("projup", om.projection_upsample)


def idwt_block(sd, p, inp, skip, hf):
    ZONE.append("lf:" + p)
    low = F.conv3d(inp, sd[f"{p}.conv_lf_block.conv.weight"], padding=1)
    ZONE.pop()
    from oracle import haar
    fmt = POLICY.get("hf", "fp32")
    hf = tuple({k: Q[fmt](v) for k, v in d.items()} for d in hf)
    rec = Q[POLICY.get("lf:" + p, "fp32").split(">")[-1]](haar.waverec3((low,) + tuple(hf), "db1"))
    return om.res_block(sd, f"{p}.conv_block", torch.cat((rec, skip), 1))


om.idwt_block = idwt_block

which = sys.argv[1] if len(sys.argv) > 1 else "unit"
cfg = ModelConfig(img_size=(128,) * 3)
if which == "unit":
    sd = make_state_dict(cfg, seed=0)
else:
    sys.path.insert(0, ROOT)
    from oracle.ref_harness import import_reference  # noqa
    raise SystemExit("init weights: use the GPU probe")
x = seeded_randn((1, 4, 128, 128, 128), 1)
torch.set_grad_enabled(False)
POLICY.clear()
ref = om.waveformer_forward(sd, x, cfg)
convz = ["res:encoder1", "res:encoder2", "res:encoder3", "res:encoder4", "calib", "lf:decoder4", "lf:decoder3", "lf:decoder2",
         "res:decoder4", "res:decoder3", "res:decoder2", "res:decoder1", "projup", "top:transp"]


def run(tag, pol):
    POLICY.clear(); POLICY.update(pol)
    y = om.waveformer_forward(sd, x, cfg); e = y - ref
    same = (y.argmax(1) == ref.argmax(1)).float().mean()
    print(f"{tag:58s} max-rel {float(e.abs().max()/ref.abs().max()):.5f} rel-L2 {float(e.norm()/ref.norm()):.5f} argmax {float(same):.5f}", flush=True)


cur = {z: "bf16" for z in convz}
cur.update({"res:encoder2": "fp16", "res:encoder3": "fp16", "res:encoder4": "fp16", "attn": "fp16>fp32", "ffn": "bf16",
            "merge": "bf16>fp32", "hf": "bf16", "top:out": "bf16>fp32"})
run("r01 policy (simulated)", cur)
p = dict(cur); p.update({z: "fp16" for z in convz}); p["top:out"] = "fp16>fp32"
run("conv U-Net fp16, ffn bf16, hf bf16", p)
p2 = dict(p); p2["hf"] = "fp16"
run("conv U-Net fp16, ffn bf16, hf fp16", p2)
p3 = dict(p2); p3["ffn"] = "fp16"; p3["merge"] = "fp16>fp32"
run("conv U-Net fp16, ffn fp16, hf fp16", p3)
p4 = dict(p3); p4["attn"] = "fp32"
run("... + attention fp32", p4)
p5 = dict(p3); p5["ffn"] = "fp32"; p5["merge"] = "fp32"
run("conv fp16, hf fp16, attn fp16, ffn fp32", p5)
for z in ["res:encoder1", "res:decoder1", "res:decoder2", "projup", "top:transp"]:
    pz = dict(p3); pz[z] = "fp32"
    run(f"all-fp16 but {z} fp32", pz)

if os.environ.get("WF_ZONES"):
    print("--- single-zone attribution, fp16 (operands + result) ---")
    for z in convz + ["attn", "ffn", "merge", "hf"]:
        fmt = "fp16>fp32" if z in ("attn", "merge") else "fp16"
        run(f"only {z} {fmt}", {z: fmt})
