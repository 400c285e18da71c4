"""Re-hosted ``Predictor`` (reference ``light_training/prediction.py:28-228``): the mirror test-time augmentation around
the sliding-window inferer, the resampling of the averaged probabilities and the un-cropping, with every per-voxel step
kept on the device.

The reference moves each of the 2^k mirrored predictions to the host (``.cpu()`` at ``prediction.py:126,135-155``) and
averages there - 8 device->host copies of a 143 MB volume per case with the default ``mirror_axes=[0, 1, 2]`` - after
materialising ``torch.flip(x)`` and ``torch.flip(output)`` for every pass.  Here a mirrored pass is an index transform
inside the stitching kernels (``wf_sw_gather`` reads the mirrored window, ``wf_sw_accumulate`` scatters it back
un-mirrored, ``wf_sw_finalize`` folds the running mean into its normalisation), so no flipped copy of the input or of a
prediction is ever built, the accumulator is one fp32 device volume, and a single copy leaves the GPU.
"""
from __future__ import annotations

import itertools
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

__all__ = ["Predictor"]


class Predictor:
    def __init__(self, window_infer: Callable, mirror_axes: Optional[Sequence[int]] = None) -> None:
        self.window_infer = window_infer
        self.mirror_axes = mirror_axes

    # ------------------------------------------------------------------------------------------------- TTA ----------
    def maybe_mirror_and_predict_cuda(self, x: torch.Tensor, model, device=None, **kwargs) -> torch.Tensor:
        """Averaged prediction as a DEVICE tensor (the reference has the same method name for its on-device variant)."""
        if device is None:
            device = next(model.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("waveformer_b200.Predictor runs on CUDA only (there is no CPU fallback)")
        if hasattr(model, "to") and not hasattr(model, "_graphs"):
            model.to(device)
        x = x.to(device, non_blocking=True)
        axes = self.mirror_axes
        from .inferers import SlidingWindowInferer
        wi = self.window_infer
        fused = (isinstance(wi, SlidingWindowInferer) and wi.device is None and wi.process_group is None
                 and not (torch.distributed.is_available() and torch.distributed.is_initialized())
                 and all(int(s) >= int(r) for s, r in zip(x.shape[2:], wi.roi_size)))
        with torch.no_grad():
            if fused and axes is not None:
                assert max(axes) <= x.dim() - 3, "mirror_axes does not match the dimension of the input!"
                # same subsets in the same order as the reference (prediction.py:134-155); every mirrored pass adds
                # result / 2^k to the running mean from inside its normalisation kernel
                scale = 1.0 / (2 ** len(axes))
                pred = wi(x, model, **kwargs)
                pred.mul_(scale)
                for r in range(1, len(axes) + 1):
                    for subset in itertools.combinations(sorted(axes), r):
                        wi(x, model, flip=subset, into=(pred, scale, True), **kwargs)
                return pred
            pred = self.window_infer(x, model, **kwargs).clone()       # a foreign inferer may reuse its buffer
            if axes is not None:
                assert max(axes) <= x.dim() - 3, "mirror_axes does not match the dimension of the input!"
                # subsets in the reference's order: (0), (1), (2), (0,1), (0,2), (1,2), (0,1,2)   prediction.py:134-155
                for r in range(1, len(axes) + 1):
                    for subset in itertools.combinations(sorted(axes), r):
                        dims = tuple(a + 2 for a in subset)
                        pred += torch.flip(self.window_infer(torch.flip(x, dims), model, **kwargs), dims)
                pred /= 2 ** len(axes)
        return pred

    def maybe_mirror_and_predict(self, x: torch.Tensor, model, device=None, **kwargs) -> torch.Tensor:
        """Drop-in for ``Predictor.maybe_mirror_and_predict`` (``prediction.py:110-160``): returns a host tensor, after
        ONE device->host copy."""
        return self.maybe_mirror_and_predict_cuda(x, model, device, **kwargs).cpu()

    # -------------------------------------------------------------------------------------------- resampling -------
    @staticmethod
    def predict_raw_probability(model_output: torch.Tensor, properties) -> torch.Tensor:
        """Trilinear resampling to ``properties['shape_after_cropping_before_resample']`` into an fp16 buffer
        (``prediction.py:35-63``); all channels in one interpolate call on the device."""
        if model_output.dim() == 5:
            model_output = model_output[0]
        shape = properties["shape_after_cropping_before_resample"]
        d, w, h = (int(v) for v in shape[:3])
        with torch.no_grad():
            out = F.interpolate(model_output[None].float(), mode="trilinear", size=(d, w, h))[0]
        return out.to(torch.half)

    @staticmethod
    def labels_and_regions(probabilities: torch.Tensor):
        """argmax over channels and the BraTS region masks TC / WT / ET the reference derives from it
        (``4_predict.py:241-255``), computed on the device.  Returns ``(labels uint8 [D,H,W], regions uint8 [3,D,H,W])``."""
        labels = probabilities.argmax(dim=0).to(torch.uint8)
        tc = (labels == 1) | (labels == 3)
        wt = labels > 0
        et = labels == 3
        return labels, torch.stack([tc, wt, et]).to(torch.uint8)

    @staticmethod
    def predict_noncrop_probability(model_output, properties) -> np.ndarray:
        """Paste the prediction back into the un-cropped volume (``prediction.py:66-108``)."""
        if isinstance(model_output, torch.Tensor):
            model_output = model_output.cpu().numpy()
        sbc = [int(v.item()) if isinstance(v, torch.Tensor) else int(v) for v in properties["shape_before_cropping"][:3]]
        bbox = properties["bbox_used_for_cropping"]
        sl = tuple(slice(int(b[0]), int(b[1])) for b in bbox[:3])
        if model_output.ndim == 3:
            out = np.zeros(sbc, dtype=np.uint8)
            out[sl] = model_output
            return out
        if model_output.ndim == 4:
            out = np.zeros([model_output.shape[0]] + sbc, dtype=np.uint8)
            out[(slice(None),) + sl] = model_output
            return out
        raise ValueError("restore crop error: expected a 3-D label map or a 4-D channel stack")
