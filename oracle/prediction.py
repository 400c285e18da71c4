"""Oracle (test infrastructure only): the reference's mirror test-time augmentation and resampling, restated on the CPU.

Follows ``light_training/prediction.py``: ``Predictor.maybe_mirror_and_predict`` (``:110-160``: the plain prediction plus
one prediction per non-empty subset of ``mirror_axes``, each flipped back, averaged over ``2 ** len(mirror_axes)``) and
``Predictor.predict_raw_probability`` (``:35-63``: per-channel trilinear resampling into an fp16 buffer).
"""
from __future__ import annotations

import itertools
from typing import Callable, Optional, Sequence

import torch
import torch.nn.functional as F


def mirror_and_predict(x: torch.Tensor, window_infer: Callable[[torch.Tensor], torch.Tensor],
                       mirror_axes: Optional[Sequence[int]]) -> torch.Tensor:
    pred = window_infer(x)
    if mirror_axes is None:
        return pred
    assert max(mirror_axes) <= x.dim() - 3, "mirror_axes does not match the dimension of the input!"
    # the reference adds the subsets in the order (0), (1), (2), (0,1), (0,2), (1,2), (0,1,2)   (prediction.py:134-155)
    for r in range(1, len(mirror_axes) + 1):
        for subset in itertools.combinations(sorted(mirror_axes), r):
            dims = tuple(a + 2 for a in subset)
            pred = pred + torch.flip(window_infer(torch.flip(x, dims)), dims)
    return pred / (2 ** len(mirror_axes))


def predict_raw_probability(model_output: torch.Tensor, shape) -> torch.Tensor:
    if model_output.dim() == 5:
        model_output = model_output[0]
    d, w, h = (int(v) for v in shape)
    out = torch.zeros((model_output.shape[0], d, w, h), dtype=torch.half)
    for c in range(model_output.shape[0]):
        out[c] = F.interpolate(model_output[c][None, None].float(), mode="trilinear", size=(d, w, h))[0, 0]
    return out
