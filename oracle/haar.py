"""Oracle (test infrastructure only): 3D Haar analysis / synthesis as ``ptwt==0.1.9`` performs it.

The reference calls ``ptwt.wavedec3(x, wavelet='db1', level=L, mode='zero')`` at
``network_models/wave_helper.py:350`` and ``ptwt.waverec3(coeffs, wavelet='db1')`` at
``network_models/idwt_upsample.py:160``.  ``ptwt`` is a third-party dependency pinned at
``requirements.txt:45`` (0.1.9, with PyWavelets 1.6.0 at ``:48``) that is neither vendored under
``/root/reference`` nor installed in this image, so its published algorithm (``ptwt/conv_transform_3.py``) is
restated here.  PARITY UNPINNED at this boundary (see ``oracle/__init__.py``).

Two independent statements are given so they can check each other:

* ``wavedec3`` / ``waverec3``  - the convolution form ptwt uses: eight separable outer-product filters applied by
  ``conv3d(stride=2)`` and inverted by ``conv_transpose3d(stride=2)``.
* ``haar_cell_forward`` / ``haar_cell_inverse`` - the closed form per 2x2x2 cell,
  ``c[pqr] = (1/(2*sqrt(2))) * sum_{ijk} (-1)^(p*i+q*j+r*k) x[2z+i, 2y+j, 2x+k]`` (an 8x8 Hadamard matrix), which
  is what the CUDA kernels and ``oracle/haar3d.c`` compute.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

# Sub-band names in ptwt's order; letter order is (D, H, W) = dims (-3, -2, -1); 'a' = low-pass, 'd' = high-pass.
SUBBANDS: Tuple[str, ...] = ("aaa", "aad", "ada", "add", "daa", "dad", "dda", "ddd")
DETAIL_KEYS: Tuple[str, ...] = SUBBANDS[1:]


def _check_wavelet(wavelet: str) -> None:
    if wavelet not in ("db1", "haar"):
        raise ValueError(f"oracle restates the Haar ('db1'/'haar') transform only, got {wavelet!r}")


def haar_filter_bank(dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """[8,1,2,2,2] filter stack.

    pywt's db1 bank is dec_lo=(s,s), dec_hi=(-s,s), rec_lo=(s,s), rec_hi=(s,-s).  ptwt flips the decomposition
    filters before handing them to ``conv3d`` (a cross-correlation), giving lo=(s,s), hi=(s,-s); the reconstruction
    filters are used unflipped and are the same two vectors.  So one bank serves both directions.
    """
    s = 1.0 / math.sqrt(2.0)
    taps = {"a": torch.tensor([s, s], dtype=torch.float64), "d": torch.tensor([s, -s], dtype=torch.float64)}
    bank = []
    for name in SUBBANDS:
        hd, hh, hw = (taps[c] for c in name)
        bank.append(hd[:, None, None] * hh[None, :, None] * hw[None, None, :])
    return torch.stack(bank, 0).unsqueeze(1).to(dtype)


def _fold(x: torch.Tensor) -> Tuple[torch.Tensor, Tuple[int, ...]]:
    lead = tuple(x.shape[:-3])
    return x.reshape(-1, 1, *x.shape[-3:]), lead


def wavedec3(x: torch.Tensor, wavelet: str = "db1", level: int = 1, mode: str = "zero"):
    """``ptwt.wavedec3`` for Haar: returns ``(LL, detail_coarsest, ..., detail_finest)``; details are dicts keyed
    ``aad, ada, add, daa, dad, dda, ddd`` (insertion order), every tensor shaped like that level's LL."""
    _check_wavelet(wavelet)
    if mode not in ("zero", "constant"):
        raise ValueError("oracle restates mode='zero' only")
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"Input dtype {x.dtype} not supported")  # ptwt rejects everything but fp32/fp64
    if x.dim() < 3:
        raise ValueError("wavedec3 needs at least three dimensions")
    bank = haar_filter_bank(x.dtype).to(x.device)
    cur, lead = _fold(x)
    details = []
    for _ in range(level):
        d, h, w = cur.shape[-3:]
        # Haar: (2*filt_len-3)//2 = 0 samples of padding on both sides; an odd extent gets one zero on the right.
        if (d % 2) or (h % 2) or (w % 2):
            cur = F.pad(cur, (0, w % 2, 0, h % 2, 0, d % 2))
        y = F.conv3d(cur, bank, stride=2)  # [*, 8, d/2, h/2, w/2]
        cur = y[:, 0:1]
        details.append({k: y[:, n + 1].reshape(*lead, *y.shape[-3:]) for n, k in enumerate(DETAIL_KEYS)})
    return (cur.reshape(*lead, *cur.shape[-3:]),) + tuple(reversed(details))


def waverec3(coeffs: Sequence, wavelet: str = "db1") -> torch.Tensor:
    """``ptwt.waverec3`` for Haar: ``coeffs = (LL, detail_coarsest, ..., detail_finest)``."""
    _check_wavelet(wavelet)
    ll = coeffs[0]
    if not isinstance(ll, torch.Tensor):
        raise ValueError("First element of coeffs must be the approximation tensor")
    if ll.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"Input dtype {ll.dtype} not supported")
    bank = haar_filter_bank(ll.dtype).to(ll.device)
    cur, lead = _fold(ll)
    for lvl, det in enumerate(coeffs[1:]):
        if not isinstance(det, dict) or set(det.keys()) != set(DETAIL_KEYS):
            raise ValueError(f"Unexpected detail keys at level {lvl}: {list(det) if isinstance(det, dict) else det}")
        bands = [cur]
        for k in DETAIL_KEYS:
            t = det[k]
            if t.dtype != ll.dtype:
                raise ValueError("coefficients must share one dtype")
            t = t.reshape(-1, 1, *t.shape[-3:])
            if t.shape != cur.shape:
                # ptwt crops one trailing sample when an odd extent was padded during analysis
                if all(a - b in (0, 1) for a, b in zip(cur.shape[-3:], t.shape[-3:])) and t.shape[0] == cur.shape[0]:
                    cur = cur[..., : t.shape[-3], : t.shape[-2], : t.shape[-1]]
                    bands[0] = cur
                else:
                    raise ValueError("coefficient shape mismatch")
            bands.append(t)
        st = torch.cat(bands, 1)  # [*, 8, d, h, w]
        cur = F.conv_transpose3d(st, bank, stride=2)  # [*, 1, 2d, 2h, 2w]; Haar needs no cropping
    return cur.reshape(*lead, *cur.shape[-3:])


# ---------------------------------------------------------------------------------------------------------------
# Closed form (what the kernels compute).  hadamard8()[n, m] = (-1)^(p*i+q*j+r*k) with n=(p,q,r), m=(i,j,k).
# ---------------------------------------------------------------------------------------------------------------
def hadamard8(dtype: torch.dtype = torch.float64) -> torch.Tensor:
    h = torch.empty(8, 8, dtype=dtype)
    for n in range(8):
        p, q, r = (n >> 2) & 1, (n >> 1) & 1, n & 1
        for m in range(8):
            i, j, k = (m >> 2) & 1, (m >> 1) & 1, m & 1
            h[n, m] = -1.0 if (p * i + q * j + r * k) % 2 else 1.0
    return h


def haar_cell_forward(x: torch.Tensor) -> torch.Tensor:
    """x[..., D, H, W] (even extents) -> [..., 8, D/2, H/2, W/2] in SUBBANDS order, closed form."""
    cells = torch.stack([x[..., i::2, j::2, k::2] for i in (0, 1) for j in (0, 1) for k in (0, 1)], -4)
    h = hadamard8(x.dtype).to(x.device) * (1.0 / (2.0 * math.sqrt(2.0)))
    return torch.einsum("nm,...mdhw->...ndhw", h, cells)


def haar_cell_inverse(c: torch.Tensor) -> torch.Tensor:
    """c[..., 8, d, h, w] -> x[..., 2d, 2h, 2w]; the same orthonormal matrix (symmetric, self-inverse)."""
    h = hadamard8(c.dtype).to(c.device) * (1.0 / (2.0 * math.sqrt(2.0)))
    cells = torch.einsum("mn,...ndhw->...mdhw", h, c)
    d, hh, w = c.shape[-3:]
    out = c.new_empty(*c.shape[:-4], 2 * d, 2 * hh, 2 * w)
    m = 0
    for i in (0, 1):
        for j in (0, 1):
            for k in (0, 1):
                out[..., i::2, j::2, k::2] = cells[..., m, :, :, :]
                m += 1
    return out


def details_to_stack(ll: torch.Tensor, det: Dict[str, torch.Tensor]) -> torch.Tensor:
    return torch.stack([ll] + [det[k] for k in DETAIL_KEYS], -4)
