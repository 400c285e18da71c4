"""``MultiscaleTransformer`` - the 4-stage wavelet/attention encoder, same API as the reference
(``network_models/waveformer.py:36-334``).  Activations stay channels-last ([B, D, H, W, C]) from the patch embedding to
the stage outputs, so none of the reference's ``rearrange`` / ``permute().contiguous()`` copies exist here; stage
outputs are returned as [B, C, D, H, W] tensors with channels-last-3d strides (what the cuDNN decoder wants)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .blocks import SwinPatchEmbed
from .wave_helper import Block, PatchMerging


class MultiscaleTransformer(nn.Module):
    def __init__(self, img_size=(128, 128, 128), patch_size=2, in_chans=4, num_classes=4,
                 embed_dims=[48, 96, 192, 384], num_heads=[3, 6, 12, 24], mlp_ratios=[4, 4, 4, 4],
                 decom_levels=[3, 2, 1, 0], multi_scale_attention=True, qkv_bias=False, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, patch_norm=False, depths=[2, 2, 2, 2],
                 network_config=None):
        super().__init__()
        self.network_config = network_config or {}
        self.num_classes = num_classes
        self.depths = depths
        self.patch_norm = patch_norm
        self.patch_size = patch_size
        self.img_size = img_size
        self.levels = decom_levels
        self.multi_scale_attention = multi_scale_attention
        self.patch_embed = SwinPatchEmbed(patch_size=self.patch_size, in_chans=in_chans, embed_dim=embed_dims[0],
                                          norm_layer=norm_layer if self.patch_norm else None,
                                          spatial_dims=len(self.img_size))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        cur = 0
        for s in range(4):
            f = 2 ** (s + 1)  # the reference hard-codes img/2, img/4, img/8, img/16 (waveformer.py:111,132,153,174)
            blocks = nn.ModuleList([
                Block(dim=embed_dims[s], num_heads=num_heads[s], mlp_ratio=mlp_ratios[s], qkv_bias=qkv_bias,
                      qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[cur + i],
                      norm_layer=norm_layer, level=self.levels[s], ms_attention=self.multi_scale_attention,
                      img_size=(img_size[0] // f, img_size[1] // f, img_size[2] // f), network_config=self.network_config)
                for i in range(depths[s])])
            setattr(self, f"block{s + 1}", blocks)
            # only the LAST block's details reach the decoder (waveformer.py:287-288): skip the others' HF writes
            for blk in blocks[:-1]:
                blk.need_hf = False
            if s < 3:
                setattr(self, f"downsample_{s + 1}",
                        PatchMerging(dim=embed_dims[s], norm_layer=norm_layer, spatial_dims=len(img_size)))
            cur += depths[s]
        self.apply(self._init_weights)

    def _init_weights(self, m: nn.Module):
        cfg = self.network_config.get('initialization', {})
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=cfg.get('weight_std', 0.02))
            if m.bias is not None:
                nn.init.constant_(m.bias, cfg.get('bias_constant', 0.0))
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, cfg.get('layer_norm_bias', 0.0))
            nn.init.constant_(m.weight, cfg.get('layer_norm_weight', 1.0))
        elif isinstance(m, nn.Conv3d):
            fan_out = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels // m.groups
            m.weight.data.normal_(0, (2.0 / fan_out) ** 0.5)
            if m.bias is not None:
                m.bias.data.zero_()

    def proj_out(self, x: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[B, C, ...] -> affine-free LayerNorm over C (``waveformer.py:182-204``)."""
        if normalize:
            perm = (0,) + tuple(range(2, x.dim())) + (1,)
            inv = (0, x.dim() - 1) + tuple(range(1, x.dim() - 1))
            x = F.layer_norm(x.permute(perm), [x.shape[1]]).permute(inv)
        return x

    def init_weights(self, pretrained: str):
        if not isinstance(pretrained, str):
            raise TypeError('pretrained must be a str or None')
        state = torch.load(pretrained, map_location='cpu')
        self.load_state_dict(state['model'] if 'model' in state else state, strict=False)

    def forward_features(self, x_rgb: torch.Tensor, normalize: bool = True, stage_hook=None) -> Tuple[List[torch.Tensor], List]:
        """``stage_hook(i, out)`` (optional, inference wiring of ``Waveformer._forward_forked``) is called on the current
        stream right after stage i's output exists; it is an argument, not module state, so the forward stays re-entrant."""
        outs, outs_hf = [], []
        pe = self.patch_embed
        pe_dtype = pe.proj.weight.dtype
        t = None
        if (x_rgb.is_cuda and not torch.is_grad_enabled() and pe.norm is None and pe_dtype == torch.float32
                and tuple(pe.patch_size) == (2, 2, 2) and self.pos_drop.p == 0):
            # 4 input channels, 2^3 patches: one exact-fp32 kernel that writes the channels-last stream directly
            t = ops.patch_embed_k2s2(x_rgb, pe.proj.weight, pe.proj.bias)
        if t is None:
            t = self.pos_drop(pe(x_rgb if x_rgb.dtype == pe_dtype else x_rgb.to(pe_dtype)))
            t = t.permute(0, 2, 3, 4, 1)                                       # [B, D, H, W, C]
            if not t.is_contiguous():
                t = t.contiguous()
        for s in range(4):
            hf = ()
            for blk in getattr(self, f"block{s + 1}"):
                res = blk(t)
                t, hf = res if isinstance(res, tuple) else (res, ())
            od = getattr(self, "out_dtype", None) or t.dtype    # activation type of this output's consumer (prepare_inference)
            if isinstance(od, (list, tuple)):
                od = od[s]
            if normalize and t.is_cuda and not torch.is_grad_enabled():
                o = ops.layer_norm_cl(t, None, None, 1e-5, out_dtype=od)   # affine-free LN, stream -> decoder type
            else:
                o = F.layer_norm(t, [t.shape[-1]]) if normalize else t
                o = o if o.dtype == od else o.to(od)
            outs.append(o.permute(0, 4, 1, 2, 3))
            if stage_hook is not None:                     # start the stage's skip block on a side stream right away
                stage_hook(s, outs[-1])
            if s < 3:
                outs_hf.append(hf if hf is not None else ())
                t = getattr(self, f"downsample_{s + 1}")(t)
        return outs, outs_hf

    def forward(self, x_rgb: torch.Tensor, stage_hook=None):
        return self.forward_features(x_rgb, stage_hook=stage_hook)

    def flops(self) -> int:
        return 0
