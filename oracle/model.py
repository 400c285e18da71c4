"""Oracle (test infrastructure only): functional CPU restatement of ``Waveformer.forward``.

Plain PyTorch on the host (fp32 by default, fp64 on request), written as functions over a ``state_dict`` instead of
``nn.Module``s.  Every function cites the reference lines it follows.  Pinned against the unmodified reference by
``tests/golden/`` (see ``scripts/make_golden.py``) and, in the authoring container, by
``tests/test_oracle_vs_reference.py``.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import haar
from .state import ModelConfig

SD = Mapping[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------ attention ----
def window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    """``Block.window_partition`` (``wave_helper.py:450-461``): [B,D,H,W,C] -> [B*nW, ws^3, C]; windows ordered
    (b, zblk, yblk, xblk), tokens ordered (dz, dy, dx)."""
    b, d, h, w, c = x.shape
    x = x.reshape(b, d // ws, ws, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws * ws * ws, c)


def relative_position_bias(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``attention.py:93-95``: gather -> [heads, N, N]."""
    n = index.shape[0]
    return table[index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1)


def window_attention(sd: SD, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """``Attention.forward`` (``attention.py:83-104``) on x[B_, N, C]; dropout p=0 is the identity."""
    b_, n, c = x.shape
    hd = c // heads
    scale = hd ** -0.5  # attention.py:24 (qk_scale is None on this path)
    qkv = F.linear(x, sd[f"{p}.qkv.weight"], sd[f"{p}.qkv.bias"])
    qkv = qkv.reshape(b_, n, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * scale, qkv[1], qkv[2]
    s = q @ k.transpose(-2, -1)
    s = s + relative_position_bias(sd[f"{p}.relative_position_bias_table"], sd[f"{p}.relative_position_index"])[None]
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(o, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])


# ---------------------------------------------------------------------------------------------------- block ----
def ccf_ffn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """``CCF_FFN.forward`` (``wave_helper.py:260-294``); NOTE it returns ``x + ffn(x)`` (its own residual)."""
    b, d, h, w, c = x.shape
    n = d * h * w
    hid = sd[f"{p}.pwconv.weight"].shape[0]
    xc = x.permute(0, 4, 1, 2, 3)
    t = F.conv3d(xc, sd[f"{p}.pwconv.weight"], sd[f"{p}.pwconv.bias"])
    t = t.reshape(b, hid, n).permute(0, 2, 1)
    t = F.gelu(F.layer_norm(t, (hid,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5))
    t = t.permute(0, 2, 1).reshape(b, hid, d, h, w)
    t = F.conv3d(t, sd[f"{p}.dwconv.weight"], sd[f"{p}.dwconv.bias"], padding=1, groups=hid)
    t = t.reshape(b, hid, n).permute(0, 2, 1)
    t = F.gelu(F.layer_norm(t, (hid,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5))
    t = F.linear(t, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"]).reshape(b, d, h, w, c)
    return x + t


def block(sd: SD, p: str, x: torch.Tensor, heads: int, level: int, ws: int):
    """``Block.multi_scale_forward`` (``wave_helper.py:470-512``), eval mode (DropPath = identity).

    Returns ``(x, hf)`` with ``hf`` = tuple of detail dicts, coarsest first (``()`` for a level-0 block).
    The "window reverse" at ``:498-499`` is a plain reshape - no inverse permute - and is restated as such.
    """
    b, d, h, w, c = x.shape
    shortcut = x
    cur = F.layer_norm(x, (c,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-6)
    fused = 0
    hfs: List[Dict[str, torch.Tensor]] = []
    for _ in range(max(level, 1)):
        if level > 0:
            coeffs = haar.wavedec3(cur.permute(0, 4, 1, 2, 3).contiguous(), "db1", level=1, mode="zero")
            cur = coeffs[0].permute(0, 2, 3, 4, 1).contiguous()
            hfs.extend(coeffs[1:])
        d1, h1, w1 = cur.shape[1:4]
        a = window_attention(sd, f"{p}.attn", window_partition(cur, ws), heads)
        a = a.reshape(b, d1, h1, w1, c).permute(0, 4, 1, 2, 3)  # reshape-only reverse, then BCDHW
        if level > 0:
            fused = fused + F.interpolate(a, size=(d, h, w), mode="trilinear")  # align_corners=False
        else:
            fused = fused + a
    y = shortcut + fused.permute(0, 2, 3, 4, 1)
    n2 = F.layer_norm(y, (c,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-6)
    y = y + ccf_ffn(sd, f"{p}.mlp", n2)
    return y, tuple(reversed(hfs))


def patch_merging(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """``PatchMerging.forward`` (``wave_helper.py:170-194``) incl. the duplicated octants (x5 == x2, x6 == x3)."""
    sel = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1))
    cat = torch.cat([x[:, i::2, j::2, k::2, :] for i, j, k in sel], -1)
    c8 = cat.shape[-1]
    cat = F.layer_norm(cat, (c8,), sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    return F.linear(cat, sd[f"{p}.reduction.weight"])


def encoder(sd: SD, x: torch.Tensor, cfg: ModelConfig):
    """``MultiscaleTransformer.forward_features`` (``waveformer.py:260-322``)."""
    e = "waveformer_encoder"
    ps = cfg.patch_size
    pad = [0, (-x.shape[4]) % ps, 0, (-x.shape[3]) % ps, 0, (-x.shape[2]) % ps]  # patchembedding.py:199-205
    if any(pad):
        x = F.pad(x, pad)
    t = F.conv3d(x, sd[f"{e}.patch_embed.proj.weight"], sd[f"{e}.patch_embed.proj.bias"], stride=ps)
    t = t.permute(0, 2, 3, 4, 1)
    outs, outs_hf = [], []
    for s in range(4):
        hf = ()
        for bi in range(cfg.depths[s]):
            t, hf = block(sd, f"{e}.block{s + 1}.{bi}", t, cfg.num_heads[s], cfg.decom_levels[s], cfg.window_size(s))
        # proj_out (waveformer.py:182-204): affine-free LayerNorm over channels, default eps
        outs.append(F.layer_norm(t, (t.shape[-1],)).permute(0, 4, 1, 2, 3))
        if s < 3:
            outs_hf.append(hf)  # only the LAST block's details survive (waveformer.py:287-288)
            t = patch_merging(sd, f"{e}.downsample_{s + 1}", t)
    return outs, outs_hf


# -------------------------------------------------------------------------------------------------- decoder ----
def _inorm(x: torch.Tensor) -> torch.Tensor:
    return F.instance_norm(x, eps=1e-5)  # InstanceNorm3d(affine=False, track_running_stats=False)


def res_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """MONAI ``UnetResBlock.forward`` (``monai/networks/blocks/dynunet_block.py:98-111``)."""
    out = F.conv3d(x, sd[f"{p}.conv1.conv.weight"], padding=1)
    out = F.leaky_relu(_inorm(out), 0.01)
    out = _inorm(F.conv3d(out, sd[f"{p}.conv2.conv.weight"], padding=1))
    res = x
    if f"{p}.conv3.conv.weight" in sd:
        res = _inorm(F.conv3d(x, sd[f"{p}.conv3.conv.weight"]))
    return F.leaky_relu(out + res, 0.01)


def channel_calibration(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """``ChannelCalibration.forward`` (``network_backbone.py:103-128``)."""
    ident = F.conv3d(x, sd[f"{p}.residual.weight"], sd[f"{p}.residual.bias"])
    t = F.relu(_inorm(F.conv3d(x, sd[f"{p}.reduce.weight"], sd[f"{p}.reduce.bias"])))
    t = F.relu(_inorm(F.conv3d(t, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)))
    t = _inorm(F.conv3d(t, sd[f"{p}.expand.weight"], sd[f"{p}.expand.bias"]))
    se = t.mean(dim=(2, 3, 4))
    se = F.relu(F.linear(se, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"]))
    se = torch.sigmoid(F.linear(se, sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"]))
    return F.relu(t * se[:, :, None, None, None] + ident)


def idwt_block(sd: SD, p: str, inp: torch.Tensor, skip: torch.Tensor, hf: Sequence[Dict[str, torch.Tensor]]):
    """``UnetrIDWTBlock.forward`` (``idwt_upsample.py:138-166``) with ``hf_refinement=False`` (the default path)."""
    low = F.conv3d(inp, sd[f"{p}.conv_lf_block.conv.weight"], padding=1)
    rec = haar.waverec3((low,) + tuple(hf), "db1")
    return res_block(sd, f"{p}.conv_block", torch.cat((rec, skip), 1))


def projection_upsample(sd: SD, p: str, x: torch.Tensor, stride: int) -> torch.Tensor:
    """``ProjectionUpsample.forward`` (``wave_helper.py:71-81``)."""
    cin = x.shape[1]
    up = F.interpolate(x, scale_factor=float(stride), mode="trilinear", align_corners=True)
    t = F.conv3d(up, sd[f"{p}.conv1.1.weight"], sd[f"{p}.conv1.1.bias"], padding=1, groups=cin)
    t = F.group_norm(t, cin, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-5)
    t = F.gelu(F.conv3d(t, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"]))
    if f"{p}.conv3.0.weight" in sd:
        t = F.gelu(F.conv3d(t, sd[f"{p}.conv3.0.weight"], sd[f"{p}.conv3.0.bias"]))
        t = F.conv3d(t, sd[f"{p}.conv3.2.weight"], sd[f"{p}.conv3.2.bias"])
    else:
        t = F.conv3d(t, sd[f"{p}.conv3.weight"], sd[f"{p}.conv3.bias"])
    return t + F.conv3d(up, sd[f"{p}.res_conv.1.weight"], sd[f"{p}.res_conv.1.bias"])


def waveformer_forward(sd: SD, x: torch.Tensor, cfg: ModelConfig, return_intermediates: bool = False):
    """``Waveformer.forward`` (``network_backbone.py:380-407``), eval mode."""
    outs, outs_hf = encoder(sd, x, cfg)
    enc0 = res_block(sd, "encoder1.layer", x)
    enc1 = res_block(sd, "encoder2.layer", outs[0])
    enc2 = res_block(sd, "encoder3.layer", outs[1])
    enc3 = res_block(sd, "encoder4.layer", outs[2])
    dec5 = channel_calibration(sd, "encoder10", outs[3])
    dec4 = idwt_block(sd, "decoder4", dec5, enc3, outs_hf[-1])
    dec3 = idwt_block(sd, "decoder3", dec5, enc2, outs_hf[-2])
    dec2 = idwt_block(sd, "decoder2", dec5, enc1, outs_hf[-3])
    up4 = projection_upsample(sd, "learnable_up4", dec4, 4)
    up3 = projection_upsample(sd, "learnable_up3", dec3, 2)
    comb = torch.cat((up4, up3, dec2), 1)
    t = F.conv_transpose3d(comb, sd["decoder1.transp_conv.conv.weight"], stride=2)  # unetr_block.py:81-86
    dec1 = res_block(sd, "decoder1.conv_block", torch.cat((t, enc0), 1))
    logits = F.conv3d(dec1, sd["out.conv.conv.weight"], sd["out.conv.conv.bias"])
    if return_intermediates:
        return logits, dict(outs=outs, outs_hf=outs_hf, enc=[enc0, enc1, enc2, enc3], dec5=dec5,
                            dec=[dec4, dec3, dec2], up=[up4, up3], dec1=dec1)
    return logits


def cast_state(sd: SD, dtype: torch.dtype) -> Dict[str, torch.Tensor]:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
