#!/usr/bin/env python
"""Golden vectors for waveformer_b200.losses.DiceCELoss: the UNMODIFIED reference's loss (vendored MONAI, constructed as
at 3_train.py:72) on seeded logits / labels -> tests/golden/dice_ce_loss.npz (loss values and gradient samples).
Run in the authoring container (needs /root/reference)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import seeded_randn  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

rh.load_reference()
from monai.losses import DiceCELoss  # noqa: E402  (the vendored tree under /root/reference)

out = {}
cases = {"default": dict(to_onehot_y=True, softmax=True),
         "nobg_sq": dict(to_onehot_y=True, softmax=True, include_background=False, squared_pred=True, lambda_dice=0.7, lambda_ce=1.3),
         "jaccard_batch": dict(to_onehot_y=True, softmax=True, jaccard=True, batch=True)}
for name, kw in cases.items():
    x = (seeded_randn((2, 4, 12, 10, 14), 40) * 2.0).requires_grad_(True)
    y = torch.randint(0, 4, (2, 1, 12, 10, 14), generator=torch.Generator().manual_seed(41))
    loss = DiceCELoss(**kw)(x, y)
    loss.backward()
    out[f"{name}_loss"] = np.float64(loss.item())
    out[f"{name}_grad"] = x.grad.numpy().copy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "dice_ce_loss.npz"), **out)
print({k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items()})
