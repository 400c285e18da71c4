"""Inference precision policy (``prepare_inference``).

The reference runs fp32 everywhere (autocast is forced off, ``light_training/prediction.py:124``).  The B200 path's 16-bit
mode (``dtype=torch.bfloat16``, the name BASELINE uses for it) stores activations and weights of the convolutional U-Net in
16 bits - fp16 by default - but keeps the parts whose rounding dominates the logit error at higher precision.  The choices are
backed by attribution studies on the oracle (``scripts/precision_zones.py``, ``precision_policy_r02.py``,
``precision_policy2_r02.py``) and by measurements on B200 (``scripts/precision_probe.py``):

* the encoder's residual stream (patch embedding output, block outputs, patch merging) stays fp32; LayerNorm reads it
  in fp32 and writes the GEMM operand type, so no tensor of the stream is ever rounded to 16 bits;
* the patch embedding (4 -> 48 channels, 2^3 kernel: 0.2 % of the FLOPs) runs in fp32 on the fp32 input window - rounding
  the raw image to bf16 alone costs 4e-2 max-relative logit error;
* ``storage="fp16"`` (default): every 16-bit tensor - conv activations / weights, FFN operands and intermediates, detail
  bands, concat buffers - is fp16.  All of them are InstanceNorm'd / LayerNorm'd or one GEMM away from it, so range is no
  concern; the 10-bit mantissa cuts the logit error 3x against bf16 storage at identical tensor-core speed;
* ``attention="fp16x2"`` (default): window attention on error-compensated fp16 pairs (hi + lo for x, the weights, q, k and O;
  three tcgen05.mma per product) - the scores the softmax exponentiates are exact to fp32 level.  With plain fp16 operands
  the attention alone accounts for 4.5e-3 of a 6.2e-3 relative L2 logit error on the stress weights, with bf16 for 7e-2;
* fp32 accumulation and fp32 statistics in every normalisation, fp32 logits from the fused output head.

Measured argmax agreement with the fp32 reference: 99.955 % on the constructors' own initialisation, 99.893 % on the unit-gain
stress weights of the test suite (round 1's bf16-storage policy: 99.65 % / 99.53 %).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .network_models.attention import Attention
from .network_models.blocks import SwinPatchEmbed
from .network_models.wave_helper import Block, CCF_FFN, PatchMergingV2, ProjectionUpsample
from .network_models.waveformer import MultiscaleTransformer

__all__ = ["prepare_inference"]

# every attribute prepare_inference may leave on a module (instance attributes only; cleared before each preparation)
_POLICY_ATTRS = ("compute_dtype", "out_dtype", "hf_dtype", "skip_dtype", "tf32", "logits_dtype", "io_dtype", "split_operands", "split_input")


def prepare_inference(model: nn.Module, dtype: torch.dtype = torch.bfloat16, *, attention: str = "fp16x2",
                      fp32_stream: bool = True, skip_blocks: str = "fp16", storage: str = "fp16",
                      skip_split_input: bool = True) -> nn.Module:
    """Put ``model`` (a ``Waveformer``) into inference form on its current device.

    ``dtype=torch.float32``: nothing is rounded (parity mode, <= 1e-4 against the reference).
    ``dtype=torch.bfloat16``: the 16-bit policy in the module docstring.  ``storage`` is the 16-bit format of activations,
    weights and tensor-core operands outside attention: "fp16" (default; 10-bit mantissa at the speed of bf16 - every such
    tensor is InstanceNorm'd / LayerNorm'd or one GEMM away from it, so fp16's range is no concern) or "bf16" (round 1's
    policy: 3x the logit error, 0.3 % more argmax flips).  ``attention`` selects the operand format of the
    window-attention GEMMs: "fp16x2" (default: error-compensated fp16 pairs on the tensor cores - the scores the softmax
    exponentiates are exact to fp32 level), "fp16", "bf16", or "fp32" = CUDA-core kernels; ``fp32_stream=False`` gives the plain all-bf16
    model (``model.to(torch.bfloat16)``), kept for the precision study.  ``skip_blocks`` (only with ``storage="bf16"``) is the
    format of the residual blocks encoder2..4: "tf32" (fp32 storage, TF32 tensor-core convolutions), "fp16" (fp16 storage and
    operands) or "bf16" (no special treatment).  ``skip_split_input``: hand the fp32 stage outputs to those blocks as
    error-compensated fp16 pairs (conv1(hi) + conv1(lo)): their first InstanceNorm amplifies the input's rounding error ~5x.
    """
    model.eval()
    # a model may be re-prepared (bf16 -> fp32 parity run -> bf16): start from a clean slate, every time
    for m in model.modules():
        for name in _POLICY_ATTRS:
            if name in m.__dict__:
                delattr(m, name)
    if dtype == torch.float32:
        model.float()
        return model.to(memory_format=torch.channels_last_3d)
    if dtype != torch.bfloat16:
        raise ValueError("prepare_inference supports float32 and bfloat16 (= the 16-bit policy)")
    if attention not in ("fp16x2", "fp16", "bf16", "fp32"):
        raise ValueError("attention must be 'fp16x2', 'fp16', 'bf16' or 'fp32'")
    if skip_blocks not in ("tf32", "fp16", "bf16"):
        raise ValueError("skip_blocks must be 'tf32', 'fp16' or 'bf16'")
    if storage not in ("fp16", "bf16"):
        raise ValueError("storage must be 'fp16' or 'bf16'")
    if not fp32_stream:
        return model.to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
    h16 = torch.float16 if storage == "fp16" else torch.bfloat16
    if storage == "fp16":
        skip_blocks = "fp16"                           # they are ordinary members of the fp16 U-Net then
    attn_dtype = {"fp16x2": torch.float16, "fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[attention]
    keep = set()                                       # parameters that stay fp32 (never rounded to 16 bits)
    half = set()                                       # parameters stored as fp16 inside a bf16 model (skip blocks)

    def keep_fp32(mod: nn.Module) -> None:
        for t in list(mod.parameters()) + list(mod.buffers()):
            keep.add(id(t))

    for m in model.modules():
        if isinstance(m, SwinPatchEmbed):
            keep_fp32(m)                                # fp32 input window -> fp32 stream
        elif isinstance(m, Attention):
            keep_fp32(m)                                # fp32 master weights; 16-bit operand copies are cached per dtype
            m.compute_dtype = attn_dtype
            m.out_dtype = torch.float32
            m.split_operands = attention == "fp16x2"
        elif isinstance(m, (nn.LayerNorm, nn.GroupNorm)):
            keep_fp32(m)                                # the normalisation kernels read gamma / beta as fp32
        if isinstance(m, CCF_FFN):
            keep_fp32(m.dwconv)                         # depthwise stencils: the kernel takes fp32 taps
            m.compute_dtype = h16                       # 16-bit GEMM operands, fp32 stream in / out
        elif isinstance(m, ProjectionUpsample):
            keep_fp32(m.conv1[1])
        elif isinstance(m, PatchMergingV2):
            m.compute_dtype = h16
        elif isinstance(m, Block):
            m.hf_dtype = h16                            # detail bands go to the 16-bit decoder
        elif isinstance(m, MultiscaleTransformer):
            # stage outputs 0..2 feed the skip blocks below (cast on entry), the last one the 16-bit bottleneck
            m.out_dtype = [torch.float32, torch.float32, torch.float32, h16]
    # The residual blocks that turn the encoder's stage outputs into the decoder's skip connections (encoder2..4: identity
    # shortcut, 48 / 96 / 192 channels at 64^3 / 32^3 / 16^3, 5 % of the step) are the most rounding-sensitive convolutions
    # of the network (scripts/precision_policy_r02.py: in bf16 they alone are half of the logit error, in fp16 still the
    # largest single zone), because their identity shortcut carries the stage output straight into every decoder level.
    for name in ("encoder2", "encoder3", "encoder4"):
        blk = getattr(model, name, None)
        if blk is None:
            continue
        blk.io_dtype = h16
        blk.split_input = bool(skip_split_input) and skip_blocks == "fp16"
        if skip_blocks == "bf16":
            blk.skip_dtype = torch.bfloat16             # only the fp32 stage output is cast on the way in
        elif skip_blocks == "tf32":
            keep_fp32(blk)
            blk.tf32 = True
        else:
            for t in list(blk.parameters()) + list(blk.buffers()):
                half.add(id(t))
            blk.skip_dtype = torch.float16
    model.logits_dtype = torch.float32                  # the fused output head stores its fp32 accumulators
    for t in list(model.parameters()) + list(model.buffers()):
        if t.is_floating_point():
            t.data = t.data.float() if id(t) in keep else t.data.to(torch.float16 if id(t) in half else h16)
    return model.to(memory_format=torch.channels_last_3d)
