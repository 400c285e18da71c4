#!/usr/bin/env python
"""Generate ``tests/golden/*`` by running the UNMODIFIED reference (``/root/reference``) on seeded inputs.

Runs only in the authoring container (the GPU box has no reference tree).  The reference modules are imported where
they lie through ``oracle/ref_harness.py`` (four third-party imports stubbed, see there); weights come from
``oracle.state.make_state_dict`` (seeded, construction-order independent) and are loaded with ``strict=True``, so the
fixtures also pin the ``state_dict`` contract.  Inputs are regenerated from seeds by the tests; each fixture stores a
checksum of its input so RNG drift is detected rather than silently compared.

    python scripts/make_golden.py [--skip-volume]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.state import ModelConfig, make_state_dict, spec_as_json, state_spec  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def seeded_randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def checksum(t: torch.Tensor) -> float:
    return float(t.double().sum())


def sample_positions(numel: int, count: int, seed: int) -> np.ndarray:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (count,), generator=g).numpy().astype(np.int64)


def sub_state(sd, prefix):
    return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-volume", action="store_true", help="skip the 18-window 240x240x155 run (~4 min)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    nm = rh.load_reference()
    torch.set_grad_enabled(False)

    # ---- 1. state_dict contract ------------------------------------------------------------------------------
    for img in (64, 128):
        cfg = ModelConfig(img_size=(img,) * 3)
        model = nm.Waveformer(**cfg.kwargs())
        ref = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        assert ref == spec_as_json(state_spec(cfg)), "oracle.state.state_spec disagrees with the reference"
        with open(os.path.join(GOLD, f"state_dict_spec_{img}.json"), "w") as f:
            json.dump(ref, f)
        idx = model.state_dict()["waveformer_encoder.block1.0.attn.relative_position_index"]
        if img == 128:
            np.savez_compressed(os.path.join(GOLD, "relative_position_index_ws8.npz"), index=idx.numpy().astype(np.int16))

    cfg = ModelConfig(img_size=(128,) * 3)
    sd = make_state_dict(cfg, seed=0)

    # ---- 2. Attention.forward, one per stage width --------------------------------------------------------------
    att = {}
    for stage, (c, h, b_) in enumerate(((48, 3, 3), (96, 6, 2), (192, 12, 1), (384, 24, 1))):
        m = nm.Attention(c, num_heads=h, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
        m.load_state_dict(sub_state(sd, f"waveformer_encoder.block{stage + 1}.1.attn"), strict=True)
        x = seeded_randn((b_, 512, c), 100 + stage)
        y = m(x)
        att[f"in_sum_{c}"] = checksum(x)
        att[f"out_{c}"] = y.numpy()
    np.savez_compressed(os.path.join(GOLD, "attention_ws8.npz"), **att)

    # ---- 3. Block.forward (stage-1 geometry at 32^3 -> level 2 would change ws; use the true 64^3/level 3) ------
    blk = nm.Block(dim=48, num_heads=3, mlp_ratio=4, qkv_bias=True, drop_path=0.0, level=3,
                   norm_layer=lambda c: torch.nn.LayerNorm(c, eps=1e-6), img_size=(64, 64, 64)).eval()
    blk.load_state_dict(sub_state(sd, "waveformer_encoder.block1.1"), strict=True)
    x = seeded_randn((1, 64, 64, 64, 48), 200)
    y, hf = blk(x)
    pos = sample_positions(y.numel(), 8192, 201)
    out = {"in_sum": checksum(x), "pos": pos, "out": y.reshape(-1)[pos].numpy(),
           "out_sum": checksum(y), "out_abs_sum": float(y.double().abs().sum())}
    for li, d in enumerate(hf):  # coarsest first
        for key, t in d.items():
            p = sample_positions(t.numel(), 512, 300 + li)
            out[f"hf{li}_{key}"] = t.reshape(-1)[p].numpy()
            out[f"hf{li}_{key}_shape"] = np.array(t.shape)
    np.savez_compressed(os.path.join(GOLD, "block_stage1.npz"), **out)

    # ---- 4. PatchMerging (duplicated octants) --------------------------------------------------------------------
    pm = nm.PatchMerging(dim=48, norm_layer=lambda c: torch.nn.LayerNorm(c, eps=1e-6), spatial_dims=3).eval()
    pm.load_state_dict(sub_state(sd, "waveformer_encoder.downsample_1"), strict=True)
    x = seeded_randn((1, 8, 8, 8, 48), 400)
    np.savez_compressed(os.path.join(GOLD, "patch_merging.npz"), in_sum=checksum(x), out=pm(x).numpy())

    # ---- 5. UnetrIDWTBlock (decoder3 geometry: 2-level synthesis) -------------------------------------------------
    dec = nm.IDWTBlock(spatial_dims=3, in_channels=384, out_channels=96, stage=2, hf_refinement=False,
                       wavelet="db1", kernel_size=3, norm_name="instance", res_block=True).eval()
    dec.load_state_dict(sub_state(sd, "decoder3"), strict=True)
    inp = seeded_randn((1, 384, 4, 4, 4), 500)
    skip = seeded_randn((1, 96, 16, 16, 16), 501)
    keys = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")
    hf = ({k: seeded_randn((1, 96, 4, 4, 4), 510 + i) for i, k in enumerate(keys)},
          {k: seeded_randn((1, 96, 8, 8, 8), 520 + i) for i, k in enumerate(keys)})
    np.savez_compressed(os.path.join(GOLD, "idwt_block.npz"), in_sum=checksum(inp) + checksum(skip),
                        out=dec(inp, skip, hf).numpy())

    # ---- 6. Waveformer.forward, 1x4x128^3 (BASELINE config 1) ----------------------------------------------------
    model = nm.Waveformer(**cfg.kwargs()).eval()
    model.load_state_dict(sd, strict=True)
    x = seeded_randn((1, 4, 128, 128, 128), 1)
    t0 = time.time()
    y = model(x)
    print(f"reference forward 128^3: {time.time() - t0:.1f}s on {torch.get_num_threads()} threads")
    pos = sample_positions(y.numel(), 16384, 2)
    np.savez_compressed(
        os.path.join(GOLD, "waveformer_128.npz"), in_sum=checksum(x), pos=pos, logits=y.reshape(-1)[pos].numpy(),
        absmax=float(y.abs().max()), chan_mean=y.mean(dim=(0, 2, 3, 4)).numpy(), chan_std=y.std(dim=(0, 2, 3, 4)).numpy(),
        label_hist=np.bincount(y.argmax(1).reshape(-1).numpy(), minlength=4))
    outs, outs_hf = model.waveformer_encoder(x)
    enc = {}
    for i, o in enumerate(outs):
        p = sample_positions(o.numel(), 2048, 10 + i)
        enc[f"out{i}"] = o.reshape(-1)[p].numpy()
    for si, hfs in enumerate(outs_hf):
        for li, d in enumerate(hfs):
            t = d["dad"]
            enc[f"hf_s{si}_l{li}_dad"] = t.reshape(-1)[sample_positions(t.numel(), 512, 20 + 4 * si + li)].numpy()
    np.savez_compressed(os.path.join(GOLD, "encoder_128.npz"), **enc)

    # ---- 7. MONAI sliding window: stitching pinned with a cheap predictor ----------------------------------------
    Inferer = rh.load_reference_inferer()
    wconv = seeded_randn((3, 2, 3, 3, 3), 600) * 0.2

    def cheap(p):
        return torch.nn.functional.conv3d(p, wconv, padding=1)

    sw = {}
    for name, shape, roi, ov, mode, bs in (("a", (1, 2, 40, 36, 30), (16, 16, 16), 0.5, "gaussian", 2),
                                           ("b", (2, 2, 20, 33, 17), (16, 16, 16), 0.25, "gaussian", 3),
                                           ("c", (1, 2, 12, 40, 16), (16, 16, 16), 0.5, "constant", 4)):
        x = seeded_randn(shape, 610 + ord(name))
        y = Inferer(roi_size=roi, sw_batch_size=bs, overlap=ov, mode=mode)(x, cheap)
        sw[f"{name}_out"] = y.numpy()
        sw[f"{name}_in_sum"] = checksum(x)
    np.savez_compressed(os.path.join(GOLD, "sliding_window_small.npz"), **sw)
    from monai.data.utils import compute_importance_map, dense_patch_slices
    from monai.inferers.utils import _get_scan_interval
    iv = _get_scan_interval((240, 240, 155), (128, 128, 128), 3, (0.5, 0.5, 0.5))
    sl = dense_patch_slices((240, 240, 155), (128, 128, 128), iv)
    imap = compute_importance_map((128, 128, 128), mode="gaussian", sigma_scale=0.125)
    with open(os.path.join(GOLD, "windows_240x240x155.json"), "w") as f:
        json.dump({"interval": list(iv), "starts": [[int(s.start) for s in w] for w in sl],
                   "imap_min": float(imap.min()), "imap_max": float(imap.max()), "imap_sum": float(imap.double().sum()),
                   "imap_diag": [float(imap[i, i, i]) for i in range(0, 128, 8)]}, f)

    # ---- 8. BASELINE config 3: 1x4x240x240x155 through the reference inferer + reference model (CPU, minutes) ----
    if not args.skip_volume:
        x = seeded_randn((1, 4, 240, 240, 155), 0)
        inferer = Inferer(roi_size=(128, 128, 128), sw_batch_size=2, overlap=0.5, mode="gaussian")
        t0 = time.time()
        y = inferer(x, model)
        dt = time.time() - t0
        print(f"reference sliding window 240x240x155: {dt:.1f}s on {torch.get_num_threads()} threads")
        pos = sample_positions(y.numel(), 32768, 3)
        np.savez_compressed(
            os.path.join(GOLD, "volume_240x240x155.npz"), in_sum=checksum(x), pos=pos, logits=y.reshape(-1)[pos].numpy(),
            absmax=float(y.abs().max()), label_hist=np.bincount(y.argmax(1).reshape(-1).numpy(), minlength=4),
            cpu_seconds=dt, cpu_threads=torch.get_num_threads())
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
