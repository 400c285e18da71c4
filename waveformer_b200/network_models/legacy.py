"""Names the reference exports from ``network_models`` but never reaches from ``Waveformer.forward`` (SURVEY.md 2.1:
"dead code also exported").  Kept as small PyTorch modules so ``from network_models import ...`` keeps working; they
are not part of the accelerated path."""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class DWConv(nn.Module):
    """2D depthwise conv on [B, N, C] tokens (``wave_helper.py:87-121``)."""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)
        self.dim = dim

    def forward(self, x, H, W):
        B, N, C = x.shape
        y = self.dwconv(x.transpose(1, 2).reshape(B, C, H, W))
        return y.flatten(2).transpose(1, 2)


class Mlp(nn.Module):
    """fc1 -> act -> drop -> fc2 -> drop (``wave_helper.py:302-341``)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        for m in (self.fc1, self.fc2):
            nn.init.trunc_normal_(m.weight, std=.02)
            nn.init.zeros_(m.bias)

    def forward(self, x, H=None, W=None):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class OverlapPatchEmbed(nn.Module):
    """Overlapping 2D patch embedding (``wave_helper.py:571-614``)."""

    def __init__(self, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        patch_size = to_2tuple(patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                              padding=(patch_size[0] // 2, patch_size[1] // 2))
        self.norm = nn.LayerNorm(embed_dim)

    def forward(self, x):
        x = self.proj(x)
        _, _, H, W = x.shape
        return self.norm(x.flatten(2).transpose(1, 2)), H, W


class PatchEmbed(nn.Module):
    """3D patch embedding returning tokens (``wave_helper.py:616-690``)."""

    def __init__(self, img_size=(96, 96, 96), patch_size=2, in_chans=1, embed_dim=48, use_conv_embed=False,
                 norm_layer=None, use_pre_norm=False, is_stem=False):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.patches_resolution = [s // patch_size for s in img_size]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1] * self.patches_resolution[2]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.use_pre_norm, self.use_conv_embed = use_pre_norm, use_conv_embed
        if use_conv_embed:
            k, p, s = (7, 2, 4) if is_stem else (3, 1, 2)
            self.kernel_size = k
            self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=k, stride=s, padding=p)
        else:
            self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        if use_pre_norm:
            self.pre_norm = nn.GroupNorm(1, in_chans) if norm_layer is not None else None
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        if self.use_pre_norm and self.pre_norm is not None:
            x = self.pre_norm(x)
        x = self.proj(x)
        _, _, D, H, W = x.shape
        x = x.flatten(2).transpose(1, 2).contiguous()
        if self.norm is not None:
            x = self.norm(x)
        return x, D, H, W


class PosCNN(nn.Module):
    """Conditional positional encoding (``wave_helper.py:692-708``)."""

    def __init__(self, in_chans, embed_dim=768, s=1):
        super().__init__()
        self.proj = nn.Sequential(nn.Conv2d(in_chans, embed_dim, 3, s, 1, groups=embed_dim), nn.GELU(),
                                  nn.Conv2d(embed_dim, embed_dim, 1, 1, 0))
        self.s = s

    def forward(self, x, H, W):
        B, N, C = x.shape
        feat = x.transpose(1, 2).view(B, C, H, W)
        y = self.proj(feat) + feat if self.s == 1 else self.proj(feat)
        return y.flatten(2).transpose(1, 2)

    def no_weight_decay(self):
        return ['proj.%d.weight' % i for i in range(4)]


class ProjectionHead(nn.Module):
    """Contrastive projection head (``network_backbone.py:35-63``); BN+ReLU as ``ModuleHelper.BNReLU('torchbn')``."""

    def __init__(self, dim_in: int, proj_dim: int = 256, proj: str = 'convmlp', bn_type: str = 'torchbn'):
        super().__init__()
        if proj == 'linear':
            self.proj = nn.Conv2d(dim_in, proj_dim, kernel_size=1)
        elif proj == 'convmlp':
            if bn_type != 'torchbn':
                raise ValueError(f"bn_type {bn_type!r} is not available here")
            self.proj = nn.Sequential(nn.Conv3d(dim_in, dim_in, kernel_size=1),
                                      nn.Sequential(nn.BatchNorm3d(dim_in), nn.ReLU()),
                                      nn.Conv3d(dim_in, proj_dim, kernel_size=1))
        else:
            raise ValueError(f"Unknown projection type: {proj}")

    def forward(self, x):
        return F.normalize(self.proj(x), p=2, dim=1)
