"""Where does the 3^3 tensor-core convolution spend its clocks?  Runs the diagnostic twin (wf_conv3d_k3_c48_stage_clocks) on the
BASELINE shape (2 x 48 x 128^3, fp16) and prints, per warp role, the share of the kernel's clocks spent in each phase."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waveformer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(2, 128, 128, 128, 48, generator=g).half().to(dev).permute(0, 4, 1, 2, 3)
w = (torch.randn(48, 48, 3, 3, 3, generator=g) / 36).half().to(dev)
st = ops.instance_norm_stats(x)

clk = torch.zeros(148, 3, 8, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.conv3d_k3_c48(x, w, stage_clocks=clk)
torch.cuda.synchronize()
c = clk.double().cpu().mean(0)
names = [("producer", ["wait: free ring slot"]),
         ("issuer", ["wait: free accumulator slot", "wait: staged row (TMA)", "tcgen05.mma issue + commit"]),
         ("epilogue", ["wait: finished row", "tcgen05.ld", "zero + release", "staging tile free (barrier)", "pack + statistics + st.shared",
                       "fence + barrier + bulk store"])]
rows = 2 * 128 * 128 / 148
print(f"rolling-row kernel: {c[0, 7]:.0f} clocks per CTA, {c[0, 7] / rows:.0f} per output row")
for r, (role, phases) in enumerate(names):
    acc = 0.0
    for q, ph in enumerate(phases):
        print(f"  {role:8s} {ph:28s} {100 * c[r, q] / c[r, 7]:5.1f} %   {c[r, q] / rows:7.0f} clk per output row")
        acc += c[r, q]
    print(f"  {role:8s} {'other':28s} {100 * (c[r, 7] - acc) / c[r, 7]:5.1f} %")

clk = torch.zeros(148, 8, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.conv3d_k3_c48(x, w, in_stats=st, stage_clocks=clk)
torch.cuda.synchronize()
c = clk.double().cpu()
tot = c[:, [2, 5, 7]].mean(0)
print(f"block kernel (fused input norm): clocks per CTA (loader / issuer / epilogue view): {tot[0]:.0f} / {tot[1]:.0f} / {tot[2]:.0f}")
print(f"  loader   waits: free ring slot {100 * c[:, 0].mean() / tot[0]:5.1f} %   own cp.async copies {100 * c[:, 1].mean() / tot[0]:5.1f} %")
print(f"  issuer   waits: free accumulator {100 * c[:, 3].mean() / tot[1]:5.1f} %   staged row {100 * c[:, 4].mean() / tot[1]:5.1f} %")
print(f"  epilogue waits: finished accumulators {100 * c[:, 6].mean() / tot[2]:5.1f} %")
