"""BASELINE config 5: time one training step (forward + backward [+ AdamW]) of the product model on one GPU.

x = randn(2, 4, 128^3), labels randint(0, 4), loss = softmax cross-entropy + mean soft Dice (the DiceCELoss of the
reference's 3_train.py:72, restated with torch ops), fp32 parameters and activations (the reference's training precision; under autocast the reference's own ptwt.waverec3
rejects the mixed fp32 / bf16 sub-bands, and so does this package).
Prints one JSON line.  Not part of bench.py's contract (the headline metric is inference voxels/s).
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--optimizer", action="store_true")
args = ap.parse_args()

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waveformer_b200 import ops
from waveformer_b200.network_models import Waveformer

torch.manual_seed(0)
dev = torch.device("cuda:0")
m = Waveformer(img_size=(args.size,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4,
               feat_size=[48, 96, 192, 384], num_heads=[3, 6, 12, 24], drop_path_rate=0.1).to(dev).train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4) if args.optimizer else None
x = torch.randn(args.batch, 4, args.size, args.size, args.size, device=dev)
y = torch.randint(0, 4, (args.batch, args.size, args.size, args.size), device=dev)


def loss_of(logits):
    p = logits.float().softmax(1)
    onehot = F.one_hot(y, 4).permute(0, 4, 1, 2, 3).float()
    dice = 1 - (2 * (p * onehot).sum((2, 3, 4)) + 1e-5) / (p.sum((2, 3, 4)) + onehot.sum((2, 3, 4)) + 1e-5)
    return F.cross_entropy(logits.float(), y) + dice.mean()


def step():
    m.zero_grad(set_to_none=True)
    logits = m(x)
    loss = loss_of(logits)
    loss.backward()
    if opt is not None:
        opt.step()
    return loss


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
ops.LAUNCHES = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps({"metric": "training step (fwd + bwd%s)" % (" + AdamW" if opt else ""), "ms_per_step": round(ms, 2),
                  "samples_per_s": round(args.batch / ms * 1e3, 3), "batch": args.batch, "size": args.size,
                  "precision": "fp32", "loss": round(float(loss.detach()), 5),
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                  "own_kernel_launches_per_step": ops.LAUNCHES // args.steps}))
