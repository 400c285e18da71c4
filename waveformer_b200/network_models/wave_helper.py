"""Hot-path helper modules with the reference's names and signatures (``network_models/wave_helper.py``).

``WaveletTransform3D`` and ``Block`` route the wavelet analysis and the window attention through the sm_100a kernels
(``waveformer_b200.ops``); the remaining layers are thin PyTorch modules that keep the reference's parameter names so a
reference checkpoint loads with ``strict=True``.
"""
from __future__ import annotations

import itertools
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .attention import Attention


def fused_path(x: torch.Tensor) -> bool:
    """Inference on CUDA runs the hand-written kernels; autograd (training) keeps differentiable torch ops."""
    return x.is_cuda and not torch.is_grad_enabled()


def _ln(norm: nn.LayerNorm, x: torch.Tensor, gelu: bool = False, out_dtype=None, also_bf16: bool = False):
    return ops.layer_norm_cl(x, norm.weight, norm.bias, norm.eps, gelu=gelu, out_dtype=out_dtype, also_bf16=also_bf16)


def _linear_f32_out(a: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """bf16 operands, fp32 accumulators written out unrounded (the result joins the fp32 residual stream)."""
    k = a.shape[-1]
    y = torch.mm(a.reshape(-1, k), weight.t(), out_dtype=torch.float32)
    return y.view(a.shape[:-1] + (weight.shape[0],))


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm's ``DropPath``, imported by the reference at ``wave_helper.py:26``)."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _init_like_reference(m: nn.Module) -> None:
    """Initialisation rule shared by the reference's modules (``wave_helper.py:241-254,436-448``)."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.LayerNorm):
        nn.init.zeros_(m.bias)
        nn.init.ones_(m.weight)


class WaveletTransform3D(nn.Module):
    """Drop-in for ``WaveletTransform3D`` (``wave_helper.py:343-353``): ``forward(x[B, C, D, H, W], level)`` returns
    ``(Yl, Yh)`` with ``Yh`` a tuple (coarsest level first) of dicts keyed ``aad, ada, add, daa, dad, dda, ddd``.

    Differences from ptwt, by design: bf16 inputs are accepted as well as fp32 (ptwt: fp32/fp64 only); extents must be
    even at every level (ptwt zero-pads odd extents; the path never produces them) - odd extents raise ``ValueError``.
    """

    def __init__(self, wavelet='db1', level=5, mode='zero'):
        super().__init__()
        if wavelet not in ('db1', 'haar'):
            raise ValueError(f"waveformer_b200 implements the Haar ('db1') wavelet only, got {wavelet!r}")
        if mode not in ('zero', 'constant'):
            raise ValueError(f"waveformer_b200 implements mode='zero' only, got {mode!r}")
        self.wavelet = wavelet
        self.mode = mode

    def forward(self, x: torch.Tensor, level: int):
        details = []
        cur = x
        for _ in range(int(level)):
            cur, hf = ops.dwt3d(cur, need_hf=True)
            details.append({k: hf[i] for i, k in enumerate(ops.DETAIL_KEYS)})
        return cur, tuple(reversed(details))


class InverseWaveletTransform3D(nn.Module):
    """Counterpart of ``ptwt.waverec3`` for ``(Yl,) + Yh`` in the layout ``WaveletTransform3D`` returns."""

    def forward(self, coeffs) -> torch.Tensor:
        return waverec3(coeffs)


def _stack_details(det, like: torch.Tensor) -> torch.Tensor:
    if not isinstance(det, dict) or set(det.keys()) != set(ops.DETAIL_KEYS):
        raise ValueError(f"detail coefficients must be a dict with keys {ops.DETAIL_KEYS}")
    for k in ops.DETAIL_KEYS:
        t = det[k]
        if t.shape != like.shape or t.dtype != like.dtype:
            raise ValueError(f"detail '{k}' has shape/dtype {tuple(t.shape)}/{t.dtype}, expected {tuple(like.shape)}/{like.dtype}")
    first = det[ops.DETAIL_KEYS[0]]
    # details produced by our own analysis are slices of one [7, ...] stack: recover it without a copy
    base = first._base if first._base is not None else None
    if (base is not None and base.dim() == like.dim() + 1 and base.shape[0] == 7 and base.is_contiguous()
            and all(det[k]._base is base and det[k].data_ptr() == base[i].data_ptr() for i, k in enumerate(ops.DETAIL_KEYS))):
        return base
    return torch.stack([det[k] for k in ops.DETAIL_KEYS], 0)


def waverec3(coeffs, wavelet: str = 'db1') -> torch.Tensor:
    """Multi-level Haar synthesis of ``(LL, details_coarsest, ..., details_finest)`` for [..., D, H, W] tensors."""
    if wavelet not in ('db1', 'haar'):
        raise ValueError(f"waveformer_b200 implements the Haar ('db1') wavelet only, got {wavelet!r}")
    cur = coeffs[0]
    for det in coeffs[1:]:
        cur = ops.idwt3d(cur, _stack_details(det, cur))
    return cur


class CCF_FFN(nn.Module):
    """Convolutional channel-fusion FFN (``wave_helper.py:196-294``): 1^3 conv -> LN -> GELU -> depthwise 3^3 conv ->
    LN -> GELU -> Linear, with its own residual.  Channels-last in, channels-last out; the reference's three layout
    flips become zero-copy views because the convolutions run in channels-last-3d."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, drop=0., img_size=(48, 48, 48)):
        super().__init__()
        self.D, self.H, self.W = img_size[0], img_size[1], img_size[2]
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.C_hid = hidden_features
        self.pwconv = nn.Conv3d(in_features, hidden_features, kernel_size=1, stride=1, padding=0, bias=True)
        self.dwconv = nn.Conv3d(hidden_features, hidden_features, kernel_size=3, stride=1, padding=1, bias=True,
                                groups=hidden_features)
        self.fc = nn.Linear(hidden_features, in_features)
        self.act = act_layer()
        self.norm1 = norm_layer(hidden_features)
        self.norm2 = norm_layer(hidden_features)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        _init_like_reference(m)
        if isinstance(m, nn.Conv3d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels // m.groups
            m.weight.data.normal_(0, (2.0 / fan_out) ** 0.5)
            if m.bias is not None:
                m.bias.data.zero_()

    def ffn_fused(self, x: torch.Tensor, xop: torch.Tensor, defer_bias: bool = False):
        """The FFN branch WITHOUT its residual.  ``xop`` = x in the GEMM operand type (bf16 under the inference policy,
        where x itself is the fp32 copy); the 4C-wide intermediates live in the operand type, the result is returned in
        x's type - unrounded fp32 accumulators when x is the fp32 stream."""
        C = x.shape[-1]
        cd = xop.dtype
        t = F.linear(xop, ops.cast_cached(self.pwconv.weight, cd).view(self.C_hid, C), ops.cast_cached(self.pwconv.bias, cd))
        t = _ln(self.norm1, t, gelu=True)                                  # LayerNorm + GELU, one pass
        t = ops.dwconv3d_channels_last(t, *self._packed_dwconv())          # hand-written stencil
        t = _ln(self.norm2, t, gelu=True)
        if x.dtype == torch.float32 and cd != torch.float32:
            f = _linear_f32_out(t, ops.cast_cached(self.fc.weight, cd))
            if defer_bias:
                return f, self.fc.bias       # the caller adds the bias in its fused residual pass
            return f + ops.f32_cached(self.fc.bias)
        return F.linear(t, ops.cast_cached(self.fc.weight, cd), ops.cast_cached(self.fc.bias, cd))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, D, H, W, C = x.shape
        assert D * H * W == self.D * self.H * self.W
        if fused_path(x):
            return x + self.ffn_fused(x, x.to(getattr(self, "compute_dtype", None) or x.dtype))
        # the 1^3 convolution is a per-voxel linear map: run it on the channels-last tensor directly
        t = F.linear(x, self.pwconv.weight.view(self.C_hid, C), self.pwconv.bias)
        t = self.act(self.norm1(t))
        t = self.dwconv(t.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
        t = self.act(self.norm2(t))
        return x + self.fc(t)

    def _packed_dwconv(self):
        w = self.dwconv.weight
        key = (w._version, w.data_ptr(), w.device)
        cache = getattr(self, "_dw_cache", None)
        if cache is None or cache[0] != key:
            b = self.dwconv.bias
            cache = (key, ops.repack_depthwise_weight(w), None if b is None else b.detach().float().contiguous())
            self._dw_cache = cache
        return cache[1], cache[2]

    def flops(self):
        return 0


class PatchMergingV2(nn.Module):
    """Swin patch merging (``wave_helper.py:125-167``): the 8 octants in ``itertools.product`` order."""

    def __init__(self, dim: int, norm_layer=nn.LayerNorm, spatial_dims: int = 3) -> None:
        super().__init__()
        self.dim = dim
        n = 8 if spatial_dims == 3 else 4
        self.reduction = nn.Linear(n * dim, 2 * dim, bias=False)
        self.norm = norm_layer(n * dim)

    _octants = tuple(itertools.product(range(2), range(2), range(2)))

    def _gather(self, x):
        if x.dim() == 5:
            _, d, h, w, _ = x.shape
            if (d % 2) or (h % 2) or (w % 2):
                x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2, 0, d % 2))
            return torch.cat([x[:, i::2, j::2, k::2, :] for i, j, k in self._octants], -1)
        if x.dim() == 4:
            _, h, w, _ = x.shape
            if (h % 2) or (w % 2):
                x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
            return torch.cat([x[:, j::2, i::2, :] for i, j in itertools.product(range(2), range(2))], -1)
        raise ValueError(f"expecting 4D or 5D x, got {x.shape}.")

    def _merge(self, x):
        if fused_path(x):
            cd = getattr(self, "compute_dtype", None) or x.dtype
            n = None
            if x.dim() == 5 and isinstance(self.norm, nn.LayerNorm):
                # octant gather + LayerNorm in one kernel: the [.., 8C] concatenation is never stored
                n = ops.patch_merge_layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps, self._octants, cd)
            if n is None:
                n = _ln(self.norm, self._gather(x), out_dtype=cd)
            if x.dtype == torch.float32 and cd != torch.float32:
                return _linear_f32_out(n, ops.cast_cached(self.reduction.weight, cd))   # fp32 stream stays unrounded
            return F.linear(n, ops.cast_cached(self.reduction.weight, cd))
        return self.reduction(self.norm(self._gather(x)))

    def forward(self, x):
        return self._merge(x)


class PatchMerging(PatchMergingV2):
    """The MONAI-0.9 ordering the reference trains with (``wave_helper.py:170-194``): octants
    (0,0,0) (1,0,0) (0,1,0) (0,0,1) (1,0,1) (0,1,0) (0,0,1) (1,1,1) - two are duplicates, which is the spec."""

    _octants = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1))

    def forward(self, x):
        if x.dim() == 4:
            return super().forward(x)
        if x.dim() != 5:
            raise ValueError(f"expecting 5D x, got {x.shape}.")
        return self._merge(x)


class Block(nn.Module):
    """Transformer block of the encoder (``wave_helper.py:357-512``): LN -> per level {Haar DWT -> window attention on
    LL -> trilinear upsample} summed -> residual -> CCF_FFN.

    ``forward(x[B, D, H, W, C])`` returns ``(x, hf)`` for ``level > 0`` (``hf`` = tuple of detail dicts, coarsest
    first, each tensor [B, C, d, h, w]) and ``x`` alone for ``level == 0``, like the reference.
    """

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, level=0, ms_attention=True,
                 img_size=(48, 48, 48), network_config=None):
        super().__init__()
        self.network_config = network_config or {}
        self.dim = dim
        self.img_size = img_size
        self.mlp_ratio = mlp_ratio
        self.level = level
        self.ms_attention = ms_attention
        if self.level > 0:
            self.dwt_downsamples = WaveletTransform3D(wavelet='db1', mode='zero')
        if self.ms_attention:
            self.attn_computation_level = max(self.level, 1)
        self.window_size = self.img_size[0] // pow(2, level)
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop, window_size=self.window_size, img_size=img_size)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = CCF_FFN(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer,
                           norm_layer=lambda c: nn.LayerNorm(c), drop=drop, img_size=img_size)
        self.need_hf = True  # MultiscaleTransformer clears this on blocks whose details the reference discards
        self.apply(_init_like_reference)

    def window_partition(self, x, window_size):
        B, D, H, W, C = x.shape
        x = x.view(B, D // window_size, window_size, H // window_size, window_size, W // window_size, window_size, C)
        return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, window_size, window_size, window_size, C)

    def forward(self, x):
        if self.ms_attention:
            return self.multi_scale_forward(x)
        return self.single_scale_forward(x)

    @staticmethod
    def _details_as_dict(hf: torch.Tensor):
        # hf: [7, B, d, h, w, C] channels-last stack -> dict of [B, C, d, h, w] views (no copy)
        return {k: hf[i].permute(0, 4, 1, 2, 3) for i, k in enumerate(ops.DETAIL_KEYS)}

    def _upsample_add(self, fused, a: torch.Tensor, size):
        up = F.interpolate(a.permute(0, 4, 1, 2, 3), size=size, mode='trilinear')  # align_corners=False
        up = up.permute(0, 2, 3, 4, 1)
        return up if fused is None else fused + up

    def multi_scale_forward(self, x):
        D, H, W = self.img_size
        B, _, _, _, C = x.shape
        assert D == x.shape[1] and H == x.shape[2] and W == x.shape[3]
        if fused_path(x):
            return self._multi_scale_fused(x)
        shortcut = x
        cur = self.norm1(x)
        fused = None
        hfs = []
        for _ in range(self.attn_computation_level):
            if self.level > 0:
                cur, hf = ops.dwt3d_channels_last(cur, need_hf=self.need_hf)   # no layout flips around the DWT
                if self.need_hf:
                    hfs.append(self._details_as_dict(hf))
            a = self.attn.forward_grid(cur)   # partition + attention + reshape-only reverse
            if self.level > 0:
                fused = self._upsample_add(fused, a, (D, H, W))
            else:
                fused = a if fused is None else fused + a
        y = shortcut + self.drop_path(fused)
        y = y + self.drop_path(self.mlp(self.norm2(y)))
        if self.level > 0:
            return y, tuple(reversed(hfs))
        return y

    def _multi_scale_fused(self, x):
        """Inference: x is the residual stream (fp32 or bf16).  LN -> DWT chain stays in the stream's dtype (the detail
        bands are differences of neighbouring values, so they are computed before any rounding to bf16); attention
        outputs come back in the compute dtype and are upsampled, summed and added to the stream by one kernel."""
        D, H, W = self.img_size
        cur = _ln(self.norm1, x)
        coarse, hfs = [], []
        hf_dtype = getattr(self, "hf_dtype", None)   # decoder activation type (prepare_inference); default: the stream's
        for _ in range(self.attn_computation_level):
            if self.level > 0:
                cur, hf = ops.dwt3d_channels_last(cur, need_hf=self.need_hf, hf_dtype=hf_dtype)
                if self.need_hf:
                    hfs.append(self._details_as_dict(hf))
            coarse.append(self.attn.forward_grid(cur))
        if self.level > 0:
            y = ops.upsample_trilinear_add(coarse, (D, H, W), base=x, out_dtype=x.dtype)
        else:
            y = x + coarse[0] if len(coarse) == 1 else x + sum(coarse)
        cd = getattr(self.mlp, "compute_dtype", None)
        mlp = self.mlp
        if (cd is not None and ops.ffn_fused_supported(y, mlp.C_hid, cd) and isinstance(mlp.act, nn.GELU)
                and getattr(mlp.act, "approximate", "none") == "none" and os.environ.get("WF_FFN_FUSED", "1") != "0"):
            # stages 1 / 2: norm2 + pwconv + LN + GELU in one kernel, the stencil, LN + GELU + fc + both residuals in another
            t = ops.ffn_front(y, self.norm2, mlp.pwconv.weight, mlp.pwconv.bias, mlp.norm1, cd)
            t = ops.dwconv3d_channels_last(t, *mlp._packed_dwconv())
            y = ops.ffn_back(t, mlp.norm2, mlp.fc.weight, mlp.fc.bias, y, self.norm2)
        elif cd is not None and cd != y.dtype:
            # fp32 stream: LayerNorm writes the fp32 copy (CCF_FFN's own residual, wave_helper.py:293) and the bf16 GEMM
            # operand in one pass; y + n + ffn(n) is accumulated in fp32
            n, nop = _ln(self.norm2, y, also_bf16=cd)
            f, fb = self.mlp.ffn_fused(n, nop, defer_bias=True)
            y = ops.residual_sum(y, n, f, fb)          # y + n + ffn(n) + fc bias: one pass
        else:
            y = y + self.mlp(_ln(self.norm2, y))
        if self.level > 0:
            return y, tuple(reversed(hfs))
        return y

    def single_scale_forward(self, x):
        B, D, H, W, C = x.shape
        shortcut = x
        cur = self.norm1(x)
        hfs = []
        for _ in range(self.level):
            cur, hf = ops.dwt3d_channels_last(cur, need_hf=True)
            hfs.append(self._details_as_dict(hf))
        a = self.attn.forward_grid(cur)
        if self.level > 0:
            a = self._upsample_add(None, a, (D, H, W))
        y = shortcut + self.drop_path(a)
        y = y + self.drop_path(self.mlp(self.norm2(y)))
        if self.level > 0:
            return y, tuple(reversed(hfs))
        return y

    def flops(self):
        return 0


class ProjectionUpsample(nn.Module):
    """Learnable upsampling head (``wave_helper.py:33-81``); parameter names follow the reference's Sequentials."""

    def __init__(self, in_channels, out_channels, stride=2, residual=True, use_double_conv=False):
        super().__init__()
        self.do_res = residual
        self.stride = stride
        self.use_double_conv = use_double_conv
        self.conv1 = nn.Sequential(
            nn.Upsample(scale_factor=stride, mode='trilinear', align_corners=True),
            nn.Conv3d(in_channels, in_channels, kernel_size=3, padding=1, groups=in_channels))
        self.conv2 = nn.Conv3d(in_channels, in_channels * 2, kernel_size=1, stride=1)
        if self.use_double_conv:
            self.conv3 = nn.Sequential(nn.Conv3d(in_channels * 2, in_channels, kernel_size=1), nn.GELU(),
                                       nn.Conv3d(in_channels, out_channels, kernel_size=1))
        else:
            self.conv3 = nn.Conv3d(in_channels * 2, out_channels, kernel_size=1)
        self.norm = nn.GroupNorm(num_groups=in_channels, num_channels=in_channels)
        if self.do_res:
            self.res_conv = nn.Sequential(
                nn.Upsample(scale_factor=stride, mode='trilinear', align_corners=True),
                nn.Conv3d(in_channels, out_channels, kernel_size=1, stride=1))
        self.act = nn.GELU()

    def _packed_dwconv(self):
        w = self.conv1[1].weight
        key = (w._version, w.data_ptr(), w.device)
        cache = getattr(self, "_dw_cache", None)
        if cache is None or cache[0] != key:
            b = self.conv1[1].bias
            cache = (key, ops.repack_depthwise_weight(w), None if b is None else b.detach().float().contiguous())
            self._dw_cache = cache
        return cache[1], cache[2]

    @staticmethod
    def _pointwise(t: torch.Tensor, conv: nn.Conv3d, weight: torch.Tensor = None, bias: torch.Tensor = None):
        """1^3 convolution of a channels-last tensor as one GEMM with the bias in its epilogue (the library's channels-
        last convolution adds the bias in a second full pass)."""
        w = conv.weight.view(conv.out_channels, conv.in_channels) if weight is None else weight
        return F.linear(t, w, conv.bias if bias is None else bias)

    def _gelu(self, t: torch.Tensor) -> torch.Tensor:
        ok = isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none" and t.is_contiguous() \
            and t.numel() % 8 == 0
        return ops.gelu_(t) if ok else self.act(t)

    def _forward_fused(self, x, out_buf):
        xv = x.permute(0, 2, 3, 4, 1)                                    # channels-last view [B, d, h, w, C]
        B = xv.shape[0]
        size = tuple(int(v * self.stride) for v in x.shape[2:])
        up = ops.upsample_trilinear_add([xv], size, align_corners=True)      # shared by both branches
        # GroupNorm(num_groups = C) is a per-(sample, channel) affine map a * x + d of the depthwise result; it is folded
        # into conv2 (W' = W diag(a), b' = b + W d) instead of being applied in a pass of its own, and its statistics
        # (mean, rstd per (b, c)) come out of the depthwise kernel's epilogue
        dw, mr = ops.dwconv3d_channels_last_stats(up, *self._packed_dwconv(), eps=self.norm.eps)
        c_in, c_mid = dw.shape[-1], self.conv2.out_channels
        w2 = self.conv2.weight.view(c_mid, c_in)
        b2 = self.conv2.bias
        if b2 is not None and b2.dtype != w2.dtype:
            b2 = b2.to(w2.dtype)
        wf_, bf_ = ops.groupnorm_fold_linear(mr, ops.f32_cached(self.norm.weight), ops.f32_cached(self.norm.bias), w2, b2, B,
                                             dw.dtype)
        h = torch.empty(dw.shape[:-1] + (c_mid,), dtype=dw.dtype, device=dw.device)
        for i in range(B):      # one GEMM per sample (its own folded weights); a broadcast-bias baddbmm measured 4x slower
            torch.addmm(bf_[i], dw[i].reshape(-1, c_in), wf_[i].t(), out=h[i].view(-1, c_mid))
        if self.use_double_conv:
            h = self._pointwise(self._gelu(h), self.conv3[0])       # the Sequential's own GELU is applied by the consumer below
            last = self.conv3[2]
        else:
            last = self.conv3
        dst = out_buf if out_buf is not None else torch.empty(h.shape[:-1] + (last.out_channels,), dtype=h.dtype, device=h.device)
        if self.do_res and ops.pw_gelu_dual_supported(h, up, last.out_channels, dst):
            # last 1^3 convolution on GELU(h) + the residual projection of the shared upsampled input, written into the slice: one kernel.
            # (The commuted form - res_conv's convolution at low resolution, its 48 channels upsampled in fp32 and handed over as the
            # kernel's addend - is supported by the kernel and measured slower: 85.9 against 83.6 ms per volume.)
            ops.pw_gelu_dual(h, up, last.weight, last.bias, self.res_conv[1].weight, self.res_conv[1].bias, dst)
            return dst.permute(0, 4, 1, 2, 3)
        h = self._pointwise(self._gelu(h), last)
        if self.do_res:
            # (writing the last GEMM into the strided concat slice and accumulating the residual GEMM into it with beta = 1
            # measured slower than two dense GEMMs + this add: the library copies around a strided output)
            torch.add(h, self._pointwise(up, self.res_conv[1]), out=dst)
        else:
            dst.copy_(h)
        return dst.permute(0, 4, 1, 2, 3)

    def forward(self, x, out_buf=None):
        """``out_buf`` (inference only): channels-last [B, D, H, W, C_out] destination (a slice of decoder1's input)."""
        if fused_path(x):
            return self._forward_fused(x, out_buf)
        up = self.conv1[0](x)          # one upsample shared by both branches (the reference computes it twice)
        dw = self.conv1[1](up)
        y = self.conv3(self.act(self.conv2(self.norm(dw))))
        if self.do_res:
            y = y + self.res_conv[1](up)
        return y
