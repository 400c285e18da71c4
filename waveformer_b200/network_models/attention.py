"""``Attention`` - multiscale window self-attention on the LL band, same API as the reference
(``network_models/attention.py:15-104``), computed by the hand-written CUDA kernels behind ``wf_window_attn_fwd``."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import ops


def build_relative_position_index(ws: int) -> torch.Tensor:
    """int64 [ws^3, ws^3] buffer of ``attention.py:43-57``.  The depth offset is weighted by ``3*ws - 1`` (not
    ``(2*ws-1)**2``), so different offsets share table rows; checkpoints depend on it, so it is kept."""
    r = torch.arange(ws)
    z, y, x = torch.meshgrid(r, r, r, indexing="ij")
    lin = (z * (3 * ws - 1) + y * (2 * ws - 1) + x).flatten()
    shift = (ws - 1) * ((3 * ws - 1) + (2 * ws - 1) + 1)
    return (lin[:, None] - lin[None, :] + shift).long()


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., window_size=6,
                 img_size=(48, 48, 48)):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        self.window_size = window_size
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 3, num_heads))
        self.register_buffer("relative_position_index", build_relative_position_index(window_size))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)
        self._bias_cache = None   # (key, dense transposed fp32 bias)
        self._image_cache = None  # (key, 16-bit bias image)

    def _table_key(self):
        t = self.relative_position_bias_table
        return (t._version, t.data_ptr(), t.device, t.dtype)

    def _dense_bias(self) -> torch.Tensor:
        """fp32 transposed dense bias for the CUDA-core kernels, rebuilt when the table changes."""
        key = self._table_key()
        if self.training or self._bias_cache is None or self._bias_cache[0] != key:
            self._bias_cache = (key, ops.relpos_bias_expand(self.relative_position_bias_table.detach(),
                                                            self.relative_position_index))
        return self._bias_cache[1]

    def _bias_image(self, fmt: torch.dtype) -> torch.Tensor:
        """16-bit dense bias image for the tensor-core kernels, rebuilt when the table changes."""
        key = self._table_key() + (fmt,)
        if self.training or self._image_cache is None or self._image_cache[0] != key:
            self._image_cache = (key, ops.relpos_bias_image(self.relative_position_bias_table.detach(),
                                                            self.relative_position_index, fmt))
        return self._image_cache[1]

    def forward_grid(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, D1, H1, W1, C] channels-last LL grid -> window-major result viewed as [B, D1, H1, W1, C]
        (window partition and the reference's reshape-only reverse are part of the kernel's addressing)."""
        if self.training and (self.attn_drop.p > 0 or self.proj_drop.p > 0):
            raise NotImplementedError("attention dropout is not part of the fused kernel (the path uses p = 0)")
        # compute_dtype / out_dtype are set by waveformer_b200.prepare_inference: the GEMM operand format (bf16 or fp16
        # on the tensor cores, fp32 on CUDA cores) and the result type; x may be an fp32 residual stream either way
        cd = getattr(self, "compute_dtype", None) or x.dtype
        od = getattr(self, "out_dtype", None)
        if torch.is_autocast_enabled() and getattr(self, "compute_dtype", None) is None and x.is_cuda:
            # mixed-precision training (BASELINE configs[4]: bf16 autocast): the forward GEMMs take the autocast dtype on
            # the tensor cores and return fp32 like every autocast Linear feeding a LayerNorm'd stream; the backward
            # (ops._window_attention_backward) recomputes in fp32
            cd = torch.get_autocast_dtype("cuda")
            od = torch.float32
            if x.dtype not in (torch.float32, torch.bfloat16):
                x = x.float()
        tc = ops.window_attention_uses_tensor_cores(x.shape[1:4], self.dim, self.num_heads, self.window_size, cd)
        if os.environ.get("WF_ATTN_IMPL", "").startswith("s") or self.qkv.bias is None:
            tc = False   # tests force the CUDA-core kernels to cross-check the two device implementations
        if cd == torch.float16 and not tc:
            cd = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32   # fp16 exists as a tensor-core format only
            od = None
        if not tc and od is not None and od != cd:
            od = None
        img = self._bias_image(cd) if tc else None
        dense = None if tc else self._dense_bias()
        # split_operands (prepare_inference, attention="fp16x2"): error-compensated fp16 pairs on the tensor cores
        split = bool(getattr(self, "split_operands", False)) and tc and cd == torch.float16
        y = ops.window_attention(x, self.qkv.weight, self.qkv.bias, self.proj.weight, self.proj.bias,
                                 self.relative_position_bias_table, self.relative_position_index, dense,
                                 self.num_heads, self.window_size, self.scale, cd, img, od, split)
        want = getattr(self, "out_dtype", None) or (od if torch.is_autocast_enabled() else None)
        return y if want is None or y.dtype == want else y.to(want)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b_, n, c = x.shape
        ws = self.window_size
        if n != ws ** 3 or c != self.dim:
            raise ValueError(f"expected [B_, {ws ** 3}, {self.dim}], got {tuple(x.shape)}")
        return self.forward_grid(x.reshape(b_, ws, ws, ws, c)).reshape(b_, n, c)
