"""Convolutional building blocks of the path, with the reference's parameter names.

The reference takes these from its vendored MONAI (``monai/networks/blocks/dynunet_block.py:25-111,247-299``,
``unetr_block.py:22-86,209-269``, ``patchembedding.py:147-225``, ``convolutions.py:25-171``).  They stay library
convolutions here (cuDNN; north star limits the hand-written kernels to DWT / attention / IDWT) but are laid out so that
``state_dict()`` keys match the reference exactly (``<block>.conv.weight`` etc.) and so that they run on
channels-last-3d bf16 activations without layout copies.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


def use_fused(x: torch.Tensor) -> bool:
    """Inference on CUDA takes the fused sm_100a kernels; autograd (training) keeps the differentiable torch modules."""
    return x.is_cuda and not torch.is_grad_enabled()


def _same_padding(kernel_size: int, stride: int) -> int:
    pad = (kernel_size - stride + 1) / 2  # dynunet_block.py:301-310
    if pad < 0:
        raise AssertionError("padding value should not be negative, please change the kernel size and/or stride.")
    return int(pad)


class ConvOnly(nn.Sequential):
    """A bare (transposed) convolution registered under the name ``conv`` (MONAI ``Convolution(conv_only=True)``)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1, bias: bool = False,
                 is_transposed: bool = False):
        super().__init__()
        pad = _same_padding(kernel_size, stride)
        if is_transposed:
            out_pad = 2 * pad + stride - kernel_size  # dynunet_block.py:313-325
            conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size, stride, pad, out_pad, bias=bias)
        else:
            conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride, pad, bias=bias)
        self.add_module("conv", conv)


def get_conv_layer(spatial_dims: int, in_channels: int, out_channels: int, kernel_size: int = 3, stride: int = 1,
                   act=None, norm=None, dropout=None, bias: bool = False, conv_only: bool = True,
                   is_transposed: bool = False) -> ConvOnly:
    if spatial_dims != 3:
        raise ValueError("waveformer_b200 builds the 3D path only")
    if dropout not in (None, 0, 0.0):
        raise NotImplementedError("dropout inside conv layers is not used on this path")
    return ConvOnly(in_channels, out_channels, kernel_size, stride, bias, is_transposed)


def _norm(norm_name: Union[Tuple, str], channels: int) -> nn.Module:
    name = norm_name[0] if isinstance(norm_name, (tuple, list)) else norm_name
    if str(name).lower() != "instance":
        raise NotImplementedError(f"norm {norm_name!r}: only MONAI's 'instance' (affine-free InstanceNorm3d) is on this path")
    return nn.InstanceNorm3d(channels)


class UnetResBlock(nn.Module):
    """conv3^3 - IN - LeakyReLU(0.01) - conv3^3 - IN, plus a 1^3-conv + IN shortcut when the width changes, add, act."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 norm_name: Union[Tuple, str], act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}),
                 dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size, stride, dropout=dropout)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size, 1, dropout=dropout)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = _norm(norm_name, out_channels)
        self.norm2 = _norm(norm_name, out_channels)
        self.downsample = in_channels != out_channels or stride != 1
        if self.downsample:
            self.conv3 = get_conv_layer(spatial_dims, in_channels, out_channels, 1, stride, dropout=dropout)
            self.norm3 = _norm(norm_name, out_channels)

    def _conv2_after_norm(self, c1: torch.Tensor, s1):
        """``conv2(lrelu(norm1(c1)))`` and, when the tensor-core kernel applies, norm2's statistics of the result.
        48 -> 48 channels on rows of 128 voxels (the two 128^3 blocks): one producer / consumer tcgen05 kernel that
        normalises its input while staging it and reduces the statistics of its output in its epilogue."""
        c2 = self.conv2.conv
        if (c1.dtype in ops.HALF_TYPES and c2.weight.dtype == c1.dtype and c2.in_channels == 48 and c2.out_channels == 48
                and c2.kernel_size == (3, 3, 3) and c2.stride == (1, 1, 1) and c2.bias is None and c1.shape[-1] == 128
                and c1.stride(1) == 1 and self.norm1.eps == self.norm2.eps):
            if c1.dtype == torch.float16:
                # fp16 rows: widening a staged row costs an FMA-pipe instruction per element (bf16: a shift), which makes the
                # in-kernel normalisation (0.93 ms) slower than a separate apply pass + the plain kernel (0.18 + 0.67 ms)
                out = ops.instance_norm_act(c1, "leakyrelu", 0.01, eps=self.norm1.eps, stats=s1)
                return ops.conv3d_k3_c48(out, c2.weight, slope=0.01, eps=self.norm2.eps)
            if s1 is None:
                s1 = ops.instance_norm_stats(c1, eps=self.norm1.eps)
            return ops.conv3d_k3_c48(c1, c2.weight, in_stats=s1, slope=0.01, eps=self.norm2.eps)
        out = ops.instance_norm_act(c1, "leakyrelu", 0.01, eps=self.norm1.eps, stats=s1)
        return self.conv2(out), None

    def _c4_fused(self, inp: torch.Tensor) -> bool:
        c1 = self.conv1.conv
        return (self.downsample and c1.in_channels == 4 and c1.kernel_size == (3, 3, 3) and c1.stride == (1, 1, 1)
                and c1.weight.dtype in ops.HALF_TYPES and c1.out_channels % 8 == 0 and 2 * c1.out_channels <= 128
                and self.conv3.conv.stride == (1, 1, 1) and self.norm1.eps == self.norm3.eps
                and inp.dtype in (torch.float32, c1.weight.dtype))

    @staticmethod
    def _head_fusable(x: torch.Tensor, head) -> bool:
        """Limits of wf_instnorm_apply_head_ndhwc (instnorm.cu): at most 16 output channels, at most 32 sixteen-byte
        packets per voxel, and the (activation, logits) type pairs it is instantiated for.  Anything else takes
        InstanceNorm + activation followed by the library 1^3 convolution - same result, one more pass."""
        k, c = head[0].shape[0], x.shape[1]
        vec = 4 if x.dtype == torch.float32 else 8
        od = head[2] or torch.float32
        pairs = {(torch.float32, torch.float32), (torch.bfloat16, torch.float32), (torch.bfloat16, torch.bfloat16),
                 (torch.float16, torch.float32), (torch.float16, torch.float16)}
        return k <= 16 and c % vec == 0 and c // vec <= 32 and (x.dtype, od) in pairs

    @staticmethod
    def _head_unfused(y: torch.Tensor, head) -> torch.Tensor:
        w, b, od = head
        out = F.conv3d(y, w.to(y.dtype), None if b is None else b.to(y.dtype))
        return out if od is None or out.dtype == od else out.to(od)

    def forward(self, inp: torch.Tensor, out_buf: torch.Tensor = None, head=None, lo: torch.Tensor = None) -> torch.Tensor:
        """``lo`` (inference only): low part of an error-compensated input pair (``inp`` = hi); conv1 sees hi + lo.
        ``out_buf`` (inference only): a [B, D, H, W, C] channels-last destination - typically the skip half of a
        decoder's concatenation buffer - the block's last kernel writes into (the torch.cat copy disappears).
        ``head`` (inference only) = (weight [K, C, 1, 1, 1], bias, out_dtype) of a 1^3 convolution that is the block's only
        consumer: the last kernel then emits the K-channel result and the block's own output is never stored."""
        if use_fused(inp) and self._c4_fused(inp):
            # 4-channel input (the network's first block): conv1, the 1^3 shortcut conv3 and both InstanceNorm statistics
            # in one tcgen05 kernel (the library convolution needs 3.3 ms for this K = 108 problem)
            # the shortcut's values are a 4-term dot product per channel: only its statistics come out of the convolution kernel,
            # the last pass recomputes it from the input voxel (no 48-channel shortcut tensor is written or read back)
            w3 = self.conv3.conv.weight
            recompute = (w3.dtype == self.conv1.conv.weight.dtype and self.norm2.eps == self.norm3.eps
                         and (out_buf is None or out_buf.dtype == w3.dtype))
            if recompute and os.environ.get("WF_C4_SHORTCUT_STATS", "moments") == "kernel":     # same-box A/B: statistics from the convolution kernel
                c1, s1, c3, s3 = ops.conv3d_c4_in_stats(inp, self.conv1.conv.weight, w3, eps=self.norm1.eps, store_shortcut=False)
            elif recompute:
                # ... and its statistics follow from the 14 first / second moments of the 4-channel input, so the convolution kernel
                # carries no accumulator columns for it at all (N = 48 instead of 96)
                c1, s1, c3, _ = ops.conv3d_c4_in_stats(inp, self.conv1.conv.weight, None, eps=self.norm1.eps)
                s3 = ops.shortcut4_stats(inp, w3, eps=self.norm3.eps)
            else:
                c1, s1, c3, s3 = ops.conv3d_c4_in_stats(inp, self.conv1.conv.weight, w3, eps=self.norm1.eps)
            out, s2 = self._conv2_after_norm(c1, s1)
            if s2 is None:
                s2 = ops.instance_norm_stats(out, eps=self.norm2.eps)
            if recompute:
                return ops.instance_norm_act_shortcut4(out, inp, w3, s2, s3, "leakyrelu", 0.01, out=out_buf)
            return ops.instance_norm_act(out, "leakyrelu", 0.01, res=c3, res_norm=True, eps=self.norm2.eps, res_stats=s3,
                                         out=out_buf, stats=s2)
        if use_fused(inp):
            # InstanceNorm + LeakyReLU, and InstanceNorm (+ InstanceNorm'd shortcut) + add + LeakyReLU: one kernel each
            k1, s1 = self.conv1.conv, None
            if (lo is None and inp.dtype in ops.HALF_TYPES and k1.weight.dtype == inp.dtype and k1.in_channels == 96
                    and k1.out_channels == 48 and k1.kernel_size == (3, 3, 3) and k1.stride == (1, 1, 1) and k1.bias is None
                    and inp.shape[-1] == 128 and inp.stride(1) == 1):
                # decoder1's first convolution on the [upsampled | skip] concatenation buffer: two passes of the rolling-row
                # tensor-core kernel (48 input channels each) that also deliver norm1's statistics
                c1, s1 = ops.conv3d_k3_c96_c48(inp, k1.weight, eps=self.norm1.eps)
            else:
                c1 = self.conv1(inp)
                if lo is not None:
                    c1 = c1 + self.conv1(lo)       # conv is linear: conv(hi) + conv(lo) = conv of the unrounded input
            out, s2 = self._conv2_after_norm(c1, s1)
            if self.downsample:
                c3 = self.conv3.conv
                if c3.kernel_size == (1, 1, 1) and c3.stride == (1, 1, 1) and inp.stride(1) == 1:
                    # the 1^3 shortcut is a per-voxel linear map: one GEMM on the channels-last view
                    res = F.linear(inp.permute(0, 2, 3, 4, 1), c3.weight.view(c3.out_channels, c3.in_channels)).permute(0, 4, 1, 2, 3)
                else:
                    res = self.conv3(inp)
                if head is not None and self._head_fusable(out, head):
                    return ops.instance_norm_act_head(out, head[0], head[1], "leakyrelu", 0.01, res=res, res_norm=True,
                                                      eps=self.norm2.eps, out_dtype=head[2], stats=s2)
                y = ops.instance_norm_act(out, "leakyrelu", 0.01, res=res, res_norm=True, eps=self.norm2.eps,
                                          out=out_buf, stats=s2)
                return y if head is None else self._head_unfused(y, head)
            if head is not None and self._head_fusable(out, head):
                return ops.instance_norm_act_head(out, head[0], head[1], "leakyrelu", 0.01, res=inp, eps=self.norm2.eps,
                                                  out_dtype=head[2], stats=s2)
            y = ops.instance_norm_act(out, "leakyrelu", 0.01, res=inp, eps=self.norm2.eps, out=out_buf, stats=s2)
            return y if head is None else self._head_unfused(y, head)
        out = self.lrelu(self.norm1(self.conv1(inp)))
        out = self.norm2(self.conv2(out))
        res = self.norm3(self.conv3(inp)) if self.downsample else inp
        return self.lrelu(out + res)


class UnetBasicBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 norm_name: Union[Tuple, str], act_name=None, dropout=None):
        super().__init__()
        self.conv1 = get_conv_layer(spatial_dims, in_channels, out_channels, kernel_size, stride, dropout=dropout)
        self.conv2 = get_conv_layer(spatial_dims, out_channels, out_channels, kernel_size, 1, dropout=dropout)
        self.lrelu = nn.LeakyReLU(negative_slope=0.01, inplace=True)
        self.norm1 = _norm(norm_name, out_channels)
        self.norm2 = _norm(norm_name, out_channels)

    def forward(self, inp: torch.Tensor, out_buf: torch.Tensor = None) -> torch.Tensor:
        if use_fused(inp):
            out = ops.instance_norm_act(self.conv1(inp), "leakyrelu", 0.01, eps=self.norm1.eps)
            return ops.instance_norm_act(self.conv2(out), "leakyrelu", 0.01, eps=self.norm2.eps, out=out_buf)
        out = self.lrelu(self.norm1(self.conv1(inp)))
        return self.lrelu(self.norm2(self.conv2(out)))


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 norm_name: Union[Tuple, str], res_block: bool = False):
        super().__init__()
        cls = UnetResBlock if res_block else UnetBasicBlock
        self.layer = cls(spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name)

    def forward(self, inp: torch.Tensor, out_buf: torch.Tensor = None) -> torch.Tensor:
        if getattr(self, "tf32", False) and use_fused(inp):
            # precision policy: this block is kept in fp32 storage with TF32 tensor-core convolutions
            w = self.layer.conv1.conv.weight
            if inp.dtype != w.dtype:
                inp = inp.to(w.dtype)
            with torch.backends.cudnn.flags(enabled=True, benchmark=torch.backends.cudnn.benchmark, allow_tf32=True):
                return self.layer(inp) if out_buf is None else self.layer(inp, out_buf)
        sd = getattr(self, "skip_dtype", None)
        if sd is not None and use_fused(inp):
            # precision policy, fp16 variant: fp16 storage / operands (10-bit mantissa like TF32) at the cost of the bf16 path
            lo = None
            if (getattr(self, "split_input", False) and sd == torch.float16 and inp.dtype == torch.float32
                    and isinstance(self.layer, UnetResBlock) and not self.layer.downsample and inp.numel() % 4 == 0):
                # the fp32 stage output as an fp16 pair: conv1's InstanceNorm amplifies the input's rounding error ~5x
                inp, lo = ops.split_f16(inp)
            elif inp.dtype != sd:
                inp = inp.to(sd)
            if lo is not None:
                y = self.layer(inp, out_buf, None, lo)
            else:
                y = self.layer(inp) if out_buf is None else self.layer(inp, out_buf)
            od = getattr(self, "io_dtype", None) or torch.bfloat16      # activation type of the surrounding U-Net
            return y if out_buf is not None or y.dtype == od else y.to(od)
        return self.layer(inp) if out_buf is None else self.layer(inp, out_buf)


class UnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, kernel_size: int,
                 upsample_kernel_size: int, norm_name: Union[Tuple, str], res_block: bool = False):
        super().__init__()
        self.transp_conv = get_conv_layer(spatial_dims, in_channels, out_channels, upsample_kernel_size,
                                          upsample_kernel_size, is_transposed=True)
        cls = UnetResBlock if res_block else UnetBasicBlock
        self.conv_block = cls(spatial_dims, 2 * out_channels, out_channels, kernel_size, 1, norm_name)

    def forward(self, inp: torch.Tensor, skip: torch.Tensor, cat_buf: torch.Tensor = None, head=None) -> torch.Tensor:
        """``cat_buf`` (inference only): [B, D, H, W, 2C] channels-last buffer whose channels [C, 2C) already hold
        ``skip`` (written there by its producer); only the upsampled half is copied in."""
        tc = self.transp_conv.conv
        if cat_buf is not None:
            c = tc.out_channels
            if (inp.dtype in ops.HALF_TYPES and tc.weight.dtype == inp.dtype and tc.kernel_size == (2, 2, 2) and tc.stride == (2, 2, 2) and tc.bias is None
                    and tc.in_channels % 16 == 0 and c % 16 == 0 and tc.in_channels <= 512 and inp.stride(1) == 1):
                # kernel == stride: one tensor-core GEMM that scatters its result into the concat buffer in place
                ops.conv_transpose3d_k2s2(inp.permute(0, 2, 3, 4, 1), tc.weight, out=cat_buf[..., :c])
            else:
                cat_buf[..., :c].copy_(self.transp_conv(inp).permute(0, 2, 3, 4, 1))
            merged = cat_buf.permute(0, 4, 1, 2, 3)
            if head is not None and isinstance(self.conv_block, UnetResBlock):
                return self.conv_block(merged, None, head)
            y = self.conv_block(merged)
            return y if head is None else F.conv3d(y, head[0], head[1]).to(head[2] or y.dtype)
        return self.conv_block(torch.cat((self.transp_conv(inp), skip), dim=1))


class UnetOutBlock(nn.Module):
    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, dropout=None):
        super().__init__()
        self.conv = get_conv_layer(spatial_dims, in_channels, out_channels, 1, 1, dropout=dropout, bias=True)

    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        return self.conv(inp)


class SwinPatchEmbed(nn.Module):
    """MONAI's Swin-style ``PatchEmbed`` (``patchembedding.py:147-225``): zero-pad to a multiple of the patch, then a
    stride-``patch`` convolution; optional LayerNorm over channels."""

    def __init__(self, patch_size: Union[int, Sequence[int]] = 2, in_chans: int = 1, embed_dim: int = 48,
                 norm_layer=nn.LayerNorm, spatial_dims: int = 3):
        super().__init__()
        if spatial_dims != 3:
            raise ValueError("waveformer_b200 builds the 3D path only")
        ps = (patch_size,) * 3 if isinstance(patch_size, int) else tuple(patch_size)
        self.patch_size = ps
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=ps, stride=ps)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _, _, d, h, w = x.shape
        pad = (0, (-w) % self.patch_size[2], 0, (-h) % self.patch_size[1], 0, (-d) % self.patch_size[0])
        if any(pad):
            x = F.pad(x, pad)
        x = self.proj(x)
        if self.norm is not None:
            x = self.norm(x.permute(0, 2, 3, 4, 1)).permute(0, 4, 1, 2, 3)
        return x
