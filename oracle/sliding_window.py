"""Oracle (test infrastructure only): MONAI's sliding-window inferer, restated for the path the reference drives.

Follows ``monai/inferers/utils.py:43-321`` (``sliding_window_inference``) for the configuration
``4_predict.py:199-205`` uses - tensor-in / tensor-out predictor, no buffering, ``mode`` gaussian or constant -
plus its helpers ``_get_scan_interval`` (``inferers/utils.py:363-384``), ``dense_patch_slices``
(``monai/data/utils.py:171-211``) and ``compute_importance_map`` (``monai/data/utils.py:1088-1138``).
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> Tuple[int, ...]:
    """``_get_scan_interval``: ``int(roi * (1 - overlap))`` (>= 1), or ``roi`` when the extent equals the roi."""
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            out.append(max(int(roi * (1 - overlap)), 1))
    return tuple(out)


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], interval: Sequence[int]) -> List[Tuple[int, ...]]:
    """``dense_patch_slices``: per axis, starts ``k*interval`` up to the first window that reaches the end, the
    last one snapped back to ``size - roi``; windows enumerated in C order (first axis slowest)."""
    per_axis = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        if iv == 0:
            num = 1
        else:
            n = int(math.ceil(float(img) / iv))
            first = next((d for d in range(n) if d * iv + roi >= img), None)
            num = first + 1 if first is not None else 1
        starts = []
        for k in range(num):
            s = k * iv
            s -= max(s + roi - img, 0)
            starts.append(s)
        per_axis.append(starts)
    return list(itertools.product(*per_axis))


def importance_map(roi_size: Sequence[int], mode: str = "gaussian", sigma_scale: float = 0.125,
                   dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``compute_importance_map``: separable gaussian (sigma = sigma_scale * extent) built in fp32, clamped from
    below at ``max(min(map), 1e-3)``; constant mode = ones."""
    if mode == "constant":
        m = torch.ones(tuple(roi_size), dtype=torch.float32)
    elif mode == "gaussian":
        m = None
        for i, n in enumerate(roi_size):
            x = torch.arange(-(n - 1) / 2.0, (n - 1) / 2.0 + 1, dtype=torch.float32)
            g = torch.exp(x ** 2 / (-2 * (n * sigma_scale) ** 2))
            m = g if m is None else m.unsqueeze(-1) * g[(None,) * i]
    else:
        raise ValueError(f"unsupported mode {mode}")
    floor = max(float(m.min()), 1e-3)
    return m.clamp_(min=floor).to(dtype)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, cval: float = 0.0) -> torch.Tensor:
    """inputs [B, C, *spatial] -> [B, K, *spatial]; weighted accumulate of every window, divided by the weight sum."""
    nsp = inputs.dim() - 2
    batch = inputs.shape[0]
    orig = tuple(inputs.shape[2:])
    roi = tuple(int(r) if r and r > 0 else int(o) for r, o in zip(roi_size, orig))
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad = []
    for k in range(nsp - 1, -1, -1):  # F.pad wants the last axis first
        diff = max(roi[k] - orig[k], 0)
        pad.extend([diff // 2, diff - diff // 2])
    if any(pad):
        inputs = F.pad(inputs, pad, mode="constant", value=cval)
    starts = window_starts(size, roi, scan_interval(size, roi, overlap))
    nwin = len(starts)
    w = importance_map(roi, mode, sigma_scale, inputs.dtype)
    out = None
    count = torch.zeros((1, 1) + size, dtype=inputs.dtype)
    for st in starts:
        count[(slice(None), slice(None)) + tuple(slice(s, s + r) for s, r in zip(st, roi))] += w
    total = nwin * batch
    for g in range(0, total, sw_batch_size):
        ids = range(g, min(g + sw_batch_size, total))
        sl = [(slice(i // nwin, i // nwin + 1), slice(None)) + tuple(slice(s, s + r) for s, r in zip(starts[i % nwin], roi))
              for i in ids]
        seg = predictor(torch.cat([inputs[s] for s in sl], 0))
        if out is None:
            out = torch.zeros((batch, seg.shape[1]) + size, dtype=inputs.dtype)
        seg = seg * w
        for s, p in zip(sl, seg):
            out[s] += p
    out = out / count
    if any(pad):
        crop = []
        for sp in range(nsp):
            lo = pad[(nsp - 1 - sp) * 2]
            crop.append(slice(lo, lo + orig[sp]))
        out = out[(slice(None), slice(None)) + tuple(crop)]
    return out
