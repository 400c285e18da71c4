"""Oracle (test infrastructure only): the ``state_dict`` contract of ``Waveformer`` and a seeded weight generator.

``state_spec`` lists every key / shape / dtype the reference model registers
(``network_models/network_backbone.py:131-378`` and the modules it builds); it is checked against the real
reference in ``tests/test_oracle_vs_reference.py`` and against ``tests/golden/state_dict_spec_128.json`` (dumped
from the reference by ``scripts/make_golden.py``).

``make_state_dict`` fills that spec with deterministic values that do not depend on module construction order, so
the reference (here), the oracle and the CUDA product (on the GPU box) all see the same weights.  Scales are
"trained-like" rather than the reference's init (which zeroes every bias and gives q.k logits of ~1e-3): unit-gain
weights make the softmax, the relative-position bias and every affine term matter in the parity tests.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch

Spec = List[Tuple[str, Tuple[int, ...], str]]  # (key, shape, kind)


class ModelConfig:
    """The constructor arguments that shape the model (defaults = BASELINE config)."""

    def __init__(self, img_size=(128, 128, 128), patch_size=2, in_chans=4, out_chans=4, depths=(2, 2, 2, 2),
                 feat_size=(48, 96, 192, 384), num_heads=(3, 6, 12, 24), decom_levels=(3, 2, 1, 0),
                 mlp_ratios=(4, 4, 4, 4)):
        self.img_size = tuple(int(v) for v in img_size)
        self.patch_size = int(patch_size)
        self.in_chans = int(in_chans)
        self.out_chans = int(out_chans)
        self.depths = tuple(depths)
        self.feat_size = tuple(feat_size)
        self.num_heads = tuple(num_heads)
        self.decom_levels = tuple(decom_levels)
        self.mlp_ratios = tuple(mlp_ratios)

    def stage_grid(self, stage: int) -> Tuple[int, int, int]:
        """Token grid of encoder stage ``stage`` (0-based): img/2, img/4, ... (``waveformer.py:111,132,153,174``)."""
        f = self.patch_size * (2 ** stage)
        return tuple(v // f for v in self.img_size)

    def window_size(self, stage: int) -> int:
        """``ws = img_size[0] // 2**level`` (``wave_helper.py:400``)."""
        return self.stage_grid(stage)[0] // (2 ** self.decom_levels[stage])

    def kwargs(self) -> Dict:
        return dict(img_size=self.img_size, patch_size=self.patch_size, in_chans=self.in_chans,
                    out_chans=self.out_chans, depths=list(self.depths), feat_size=list(self.feat_size),
                    num_heads=list(self.num_heads), drop_path_rate=0.1)


def relative_position_index(ws: int) -> torch.Tensor:
    """The persistent int64 buffer of ``Attention`` (``attention.py:43-57``).

    index[i, j] = (dz_i - dz_j + ws-1) * (3*ws-1) + (dy_i - dy_j + ws-1) * (2*ws-1) + (dx_i - dx_j + ws-1);
    note the depth stride is 3*ws-1 (NOT (2*ws-1)**2), so distinct offsets collide - that is the spec.
    """
    r = torch.arange(ws)
    z, y, x = torch.meshgrid(r, r, r, indexing="ij")
    lin = (z * (3 * ws - 1) + y * (2 * ws - 1) + x).reshape(-1)
    off = (ws - 1) * ((3 * ws - 1) + (2 * ws - 1) + 1)
    return (lin[:, None] - lin[None, :] + off).to(torch.int64)


def _res_block(spec: Spec, p: str, cin: int, cout: int) -> None:
    spec.append((f"{p}.conv1.conv.weight", (cout, cin, 3, 3, 3), "conv"))
    spec.append((f"{p}.conv2.conv.weight", (cout, cout, 3, 3, 3), "conv"))
    if cin != cout:
        spec.append((f"{p}.conv3.conv.weight", (cout, cin, 1, 1, 1), "conv"))


def _proj_upsample(spec: Spec, p: str, cin: int, cout: int, double: bool) -> None:
    spec.append((f"{p}.conv1.1.weight", (cin, 1, 3, 3, 3), "conv"))
    spec.append((f"{p}.conv1.1.bias", (cin,), "bias"))
    spec.append((f"{p}.conv2.weight", (2 * cin, cin, 1, 1, 1), "conv"))
    spec.append((f"{p}.conv2.bias", (2 * cin,), "bias"))
    if double:
        spec.append((f"{p}.conv3.0.weight", (cin, 2 * cin, 1, 1, 1), "conv"))
        spec.append((f"{p}.conv3.0.bias", (cin,), "bias"))
        spec.append((f"{p}.conv3.2.weight", (cout, cin, 1, 1, 1), "conv"))
        spec.append((f"{p}.conv3.2.bias", (cout,), "bias"))
    else:
        spec.append((f"{p}.conv3.weight", (cout, 2 * cin, 1, 1, 1), "conv"))
        spec.append((f"{p}.conv3.bias", (cout,), "bias"))
    spec.append((f"{p}.norm.weight", (cin,), "gain"))
    spec.append((f"{p}.norm.bias", (cin,), "bias"))
    spec.append((f"{p}.res_conv.1.weight", (cout, cin, 1, 1, 1), "conv"))
    spec.append((f"{p}.res_conv.1.bias", (cout,), "bias"))


def state_spec(cfg: ModelConfig) -> Spec:
    spec: Spec = []
    f = cfg.feat_size
    e = "waveformer_encoder"
    ps = cfg.patch_size
    spec.append((f"{e}.patch_embed.proj.weight", (f[0], cfg.in_chans, ps, ps, ps), "conv"))
    spec.append((f"{e}.patch_embed.proj.bias", (f[0],), "bias"))
    for s in range(4):
        c, h, ws = f[s], cfg.num_heads[s], cfg.window_size(s)
        hid = int(c * cfg.mlp_ratios[s])
        for b in range(cfg.depths[s]):
            p = f"{e}.block{s + 1}.{b}"
            spec += [
                (f"{p}.norm1.weight", (c,), "gain"), (f"{p}.norm1.bias", (c,), "bias"),
                (f"{p}.attn.relative_position_bias_table", ((2 * ws - 1) ** 3, h), "table"),
                (f"{p}.attn.relative_position_index", (ws ** 3, ws ** 3), "rpi"),
                (f"{p}.attn.qkv.weight", (3 * c, c), "linear"), (f"{p}.attn.qkv.bias", (3 * c,), "bias"),
                (f"{p}.attn.proj.weight", (c, c), "linear"), (f"{p}.attn.proj.bias", (c,), "bias"),
                (f"{p}.norm2.weight", (c,), "gain"), (f"{p}.norm2.bias", (c,), "bias"),
                (f"{p}.mlp.pwconv.weight", (hid, c, 1, 1, 1), "conv"), (f"{p}.mlp.pwconv.bias", (hid,), "bias"),
                (f"{p}.mlp.dwconv.weight", (hid, 1, 3, 3, 3), "conv"), (f"{p}.mlp.dwconv.bias", (hid,), "bias"),
                (f"{p}.mlp.fc.weight", (c, hid), "linear"), (f"{p}.mlp.fc.bias", (c,), "bias"),
                (f"{p}.mlp.norm1.weight", (hid,), "gain"), (f"{p}.mlp.norm1.bias", (hid,), "bias"),
                (f"{p}.mlp.norm2.weight", (hid,), "gain"), (f"{p}.mlp.norm2.bias", (hid,), "bias"),
            ]
        if s < 3:
            p = f"{e}.downsample_{s + 1}"
            spec += [(f"{p}.reduction.weight", (2 * c, 8 * c), "linear"),
                     (f"{p}.norm.weight", (8 * c,), "gain"), (f"{p}.norm.bias", (8 * c,), "bias")]
    _res_block(spec, "encoder1.layer", cfg.in_chans, f[0])
    _res_block(spec, "encoder2.layer", f[0], f[0])
    _res_block(spec, "encoder3.layer", f[1], f[1])
    _res_block(spec, "encoder4.layer", f[2], f[2])
    c, r = f[3], f[3] // 4
    spec += [
        ("encoder10.reduce.weight", (r, c, 1, 1, 1), "conv"), ("encoder10.reduce.bias", (r,), "bias"),
        ("encoder10.conv.weight", (r, r, 3, 3, 3), "conv"), ("encoder10.conv.bias", (r,), "bias"),
        ("encoder10.expand.weight", (c, r, 1, 1, 1), "conv"), ("encoder10.expand.bias", (c,), "bias"),
        ("encoder10.fc1.weight", (r, c), "linear"), ("encoder10.fc1.bias", (r,), "bias"),
        ("encoder10.fc2.weight", (c, r), "linear"), ("encoder10.fc2.bias", (c,), "bias"),
        ("encoder10.residual.weight", (c, c, 1, 1, 1), "conv"), ("encoder10.residual.bias", (c,), "bias"),
    ]
    for name, cout in (("decoder4", f[2]), ("decoder3", f[1]), ("decoder2", f[0])):
        spec.append((f"{name}.conv_lf_block.conv.weight", (cout, f[3], 3, 3, 3), "conv"))
        _res_block(spec, f"{name}.conv_block", 2 * cout, cout)
    _proj_upsample(spec, "learnable_up4", f[2], f[0], True)
    _proj_upsample(spec, "learnable_up3", f[1], f[0], False)
    spec.append(("decoder1.transp_conv.conv.weight", (3 * f[0], f[0], 2, 2, 2), "convT"))
    _res_block(spec, "decoder1.conv_block", 2 * f[0], f[0])
    spec.append(("out.conv.conv.weight", (cfg.out_chans, f[0], 1, 1, 1), "conv"))
    spec.append(("out.conv.conv.bias", (cfg.out_chans,), "bias"))
    return spec


def make_state_dict(cfg: ModelConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic fp32 weights for ``state_spec(cfg)``; each tensor has its own generator stream."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for n, (key, shape, kind) in enumerate(state_spec(cfg)):
        if kind == "rpi":
            sd[key] = relative_position_index(round(shape[0] ** (1.0 / 3.0)))
            continue
        g = torch.Generator().manual_seed(1_000_003 * (seed + 1) + n)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if kind in ("conv", "linear"):
            fan_in = 1
            for v in shape[1:]:
                fan_in *= v
            t *= 1.0 / math.sqrt(fan_in)
        elif kind == "convT":
            # ConvTranspose3d weight is [in, out, k, k, k]; stride == kernel so each output sees `in` taps
            t *= 1.0 / math.sqrt(shape[0])
        elif kind == "bias":
            t *= 0.05
        elif kind == "gain":
            t = 1.0 + 0.1 * t
        elif kind == "table":
            t *= 0.5
        sd[key] = t
    return sd


def spec_as_json(spec: Sequence[Tuple[str, Tuple[int, ...], str]]) -> List[List]:
    return [[k, list(s)] for k, s, _ in spec]
