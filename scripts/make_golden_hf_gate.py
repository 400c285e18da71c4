#!/usr/bin/env python
"""Golden fixture for the gated decoder path: the UNMODIFIED reference ``UnetrIDWTBlock(hf_refinement=True)``
(``network_models/idwt_upsample.py:53-166``: every detail band multiplied by ``sigmoid(conv1x1(relu(IN(dwconv(x)))))`` before
synthesis) on seeded inputs and seeded weights.  Authoring container only (needs /root/reference).

    python scripts/make_golden_hf_gate.py      ->  tests/golden/idwt_block_hf_gate.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

KEYS = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")


def seeded_randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def seeded_state(module, seed):
    """Deterministic weights for every parameter of `module`, keyed by name order (independent of torch's init RNG use)."""
    sd = {}
    for i, (k, v) in enumerate(module.state_dict().items()):
        if not v.is_floating_point():
            sd[k] = v
            continue
        t = seeded_randn(tuple(v.shape), 9000 + seed * 100 + i)
        if v.dim() > 1:
            fan = int(np.prod(v.shape[1:]))
            t = t / fan ** 0.5
        elif k.endswith("weight"):
            t = 1.0 + 0.1 * t
        else:
            t = 0.05 * t
        sd[k] = t
    return sd


def main():
    nm = rh.load_reference()
    torch.set_grad_enabled(False)
    dec = nm.IDWTBlock(spatial_dims=3, in_channels=64, out_channels=16, stage=2, hf_refinement=True, wavelet="db1",
                       kernel_size=3, norm_name="instance", res_block=True).eval()
    sd = seeded_state(dec, 1)
    dec.load_state_dict(sd, strict=True)
    inp = seeded_randn((2, 64, 4, 4, 4), 600)
    skip = seeded_randn((2, 16, 16, 16, 16), 601)
    hf = ({k: seeded_randn((2, 16, 4, 4, 4), 610 + i) for i, k in enumerate(KEYS)},
          {k: seeded_randn((2, 16, 8, 8, 8), 620 + i) for i, k in enumerate(KEYS)})
    out = dec(inp, skip, hf)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "idwt_block_hf_gate.npz"), out=out.numpy(),
                        keys=np.array(list(sd.keys())), in_sum=float(inp.double().sum() + skip.double().sum()))
    print("wrote idwt_block_hf_gate.npz", tuple(out.shape), "params", len(sd))


if __name__ == "__main__":
    main()
