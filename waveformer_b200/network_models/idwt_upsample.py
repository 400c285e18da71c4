"""IDWT upsampling decoder block, same API as the reference (``network_models/idwt_upsample.py``): the multi-level Haar
synthesis, the optional high-frequency gate multiply and the concatenation with the skip run in the sm_100a synthesis
kernel (``wf_idwt3d_ndhwc``), which writes straight into the channels [0, C) of the concat buffer."""
from __future__ import annotations

from typing import Any, Dict, Sequence, Tuple, Union

import torch
import torch.nn as nn

from .. import ops
from .blocks import UnetBasicBlock, UnetResBlock, get_conv_layer
from .wave_helper import _stack_details


class HFRefinementRes(nn.Module):
    """Per-sub-band gate ``x * sigmoid(conv1x1(relu(IN(dwconv3(x)))))`` (``idwt_upsample.py:12-50``).

    ``gate(x)`` returns only the multiplier; the product with the sub-band is fused into the synthesis kernel.
    ``forward(x)`` keeps the reference semantics (returns the gated sub-band)."""

    def __init__(self, in_channels, init_alpha=0.3, network_config=None):
        super().__init__()
        self.network_config = network_config or {}
        hf_config = self.network_config.get('hf_refinement', {})
        self.conv1 = nn.Conv3d(in_channels, in_channels, kernel_size=3, padding=1, groups=in_channels, bias=True)
        self.norm = nn.InstanceNorm3d(in_channels, affine=True)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv3d(in_channels, in_channels, kernel_size=1, bias=True)
        self.sigmoid = nn.Sigmoid() if hf_config.get('use_sigmoid', True) else None

    def gate(self, x):
        g = self.conv2(self.relu(self.norm(self.conv1(x))))
        return self.sigmoid(g) if self.sigmoid is not None else g

    def forward(self, x):
        return x * self.gate(x)


class UnetrIDWTBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, stage: int, hf_refinement: bool,
                 wavelet: str, kernel_size: Union[Sequence[int], int], norm_name: Union[Tuple, str],
                 res_block: bool = False, network_config: Dict[str, Any] = None) -> None:
        super().__init__()
        if wavelet not in ('db1', 'haar'):
            raise ValueError(f"waveformer_b200 implements the Haar ('db1') wavelet only, got {wavelet!r}")
        self.network_config = network_config or {}
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.wavelet = wavelet
        self.hf_refinement = hf_refinement
        if self.hf_refinement:
            self.hf_ref = nn.ModuleList(
                [HFRefinementRes(in_channels // pow(2, stage), network_config=self.network_config) for _ in range(stage)])
        self.conv_lf_block = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size=kernel_size, stride=1,
                                            conv_only=True, is_transposed=False)
        cls = UnetResBlock if res_block else UnetBasicBlock
        self.conv_block = cls(spatial_dims, out_channels * 2, out_channels, kernel_size=kernel_size, stride=1,
                              norm_name=norm_name)

    @staticmethod
    def _channels_last_stack(det, like_ncdhw: torch.Tensor) -> torch.Tensor:
        """Detail dict of [B, C, d, h, w] tensors -> [7, B, d, h, w, C] stack (zero-copy when the dict came from
        ``Block``, whose details are permuted views of exactly such a stack)."""
        if not isinstance(det, dict) or set(det.keys()) != set(ops.DETAIL_KEYS):
            raise ValueError(f"detail coefficients must be a dict with keys {ops.DETAIL_KEYS}")
        for k in ops.DETAIL_KEYS:
            if det[k].shape != like_ncdhw.shape or det[k].dtype != like_ncdhw.dtype:
                raise ValueError(f"detail '{k}': expected shape {tuple(like_ncdhw.shape)} and dtype {like_ncdhw.dtype}, "
                                 f"got {tuple(det[k].shape)} / {det[k].dtype}")
        cl = [det[k].permute(0, 2, 3, 4, 1) for k in ops.DETAIL_KEYS]
        base = cl[0]._base
        if (base is not None and base.dim() == 6 and base.shape[0] == 7 and base.is_contiguous()
                and all(t._base is base and t.is_contiguous() and t.data_ptr() == base[i].data_ptr()
                        for i, t in enumerate(cl))):
            return base
        return torch.stack(cl, 0).contiguous()

    def forward(self, inp, skip, hf_coeffs, cat_buf=None, out_buf=None):
        """``cat_buf`` (inference only): [B, D, H, W, 2C] channels-last buffer whose channels [C, 2C) already hold ``skip``;
        ``out_buf``: channels-last destination of the block's result (a slice of the next concatenation buffer)."""
        low = self.conv_lf_block(inp)                                   # [B, C, d, h, w]
        B, C = low.shape[:2]
        cur = low.permute(0, 2, 3, 4, 1)                                # channels-last view (copy only if NCDHW-contiguous)
        n_levels = len(hf_coeffs)
        fuse_cat = not torch.is_grad_enabled()                          # write level-0 output straight into the concat buffer
        for i, det in enumerate(hf_coeffs):
            if torch.is_autocast_enabled() and isinstance(det, dict) and len(det) and cur.dtype != next(iter(det.values())).dtype:
                # mixed-precision training: conv_lf_block ran in the autocast dtype while the encoder's detail bands are
                # fp32 (they come from the fp32 LayerNorm output).  ptwt.waverec3 rejects such a mix (and the reference
                # therefore trains with autocast disabled, trainer.py:454); here the synthesis simply runs in the wider
                # type - the residual block behind it is autocast again.
                wide = torch.promote_types(cur.dtype, next(iter(det.values())).dtype)
                cur = cur.to(wide)
                det = {k: v.to(wide) for k, v in det.items()}
            like = cur.permute(0, 4, 1, 2, 3)
            stack = self._channels_last_stack(det, like)
            gate = None
            if self.hf_refinement:
                gate = torch.stack([self.hf_ref[i].gate(det[k]).permute(0, 2, 3, 4, 1) for k in ops.DETAIL_KEYS], 0)
                gate = gate.to(stack.dtype).contiguous()
            out = None
            if fuse_cat and i == n_levels - 1:
                _, d, h, w, _ = cur.shape
                cat = cat_buf if cat_buf is not None else torch.empty((B, 2 * d, 2 * h, 2 * w, 2 * C), dtype=cur.dtype,
                                                                      device=cur.device)
                out = cat[..., :C]
            cur = ops.idwt3d_channels_last(cur, stack, gate, out)
        if fuse_cat and n_levels > 0:
            if cat_buf is None:
                cat[..., C:] = skip.permute(0, 2, 3, 4, 1)
            merged = cat.permute(0, 4, 1, 2, 3)                         # [B, 2C, D, H, W], channels-last-3d strides
        else:
            merged = torch.cat((cur.permute(0, 4, 1, 2, 3), skip), dim=1)
        return self.conv_block(merged) if out_buf is None else self.conv_block(merged, out_buf)
