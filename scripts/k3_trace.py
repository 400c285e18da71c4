"""Per-step clock trace of the rolling-row convolution's issuer (diagnostic build, WF_K3_NOBULK=8): where inside its run does a
slow CTA lose its time?"""
import os
import sys

import torch

os.environ["WF_K3_NOBULK"] = str(8 | int(os.environ.get("WF_K3_EXTRA", "0")))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waveformer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(2, 128, 128, 128, 48, generator=g).half().to(dev).permute(0, 4, 1, 2, 3)
w = (torch.randn(48, 48, 3, 3, 3, generator=g) / 36).half().to(dev)
clk = torch.zeros(148 * 24 + 148 * 512, dtype=torch.int64, device=dev)
for rep in range(3):
    clk.zero_()
    ops.conv3d_k3_c48(x, w, stage_clocks=clk)
    torch.cuda.synchronize()
    c = clk.cpu()
    tot = c[: 148 * 24].reshape(148, 3, 8)[:, 1, 7]
    tr = c[148 * 24:].reshape(148, 512)
    order = torch.argsort(tot, descending=True)[:3].tolist()
    for i in sorted(set(order + [0, 1, 73, 74])):
        t = tr[i]
        nz = int((t > 0).sum()) + 1
        d = (t[1:nz] - t[: nz - 1]).clamp_min(0)
        seg = [int(d[k: k + 16].float().mean()) for k in range(0, nz - 1, 16)]
        print(f"run {rep} cta {i:3d}: total {int(tot[i]):8d}, {nz} steps; mean clocks per step in groups of 16 steps: {seg}")
    pr = c[: 148 * 24].reshape(148, 3, 8)
    for i in ():
        print(f"run {rep} cta {i:3d}: producer wait slot {int(pr[i,0,0])} total {int(pr[i,0,7])} | issuer wait acc {int(pr[i,1,0])} wait row {int(pr[i,1,1])} mma {int(pr[i,1,2])} | "
              f"epilogue wait {int(pr[i,2,0])} ld {int(pr[i,2,1])} zero {int(pr[i,2,2])} tile-free {int(pr[i,2,3])} pack {int(pr[i,2,4])} store {int(pr[i,2,5])} total {int(pr[i,2,7])}")
