"""Oracle (test infrastructure only): import the UNMODIFIED reference from ``/root/reference``.

Only usable in the authoring container (the GPU box has no ``/root/reference``); used by ``scripts/make_golden.py``
to produce the fixtures under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent).  Nothing from the reference is copied: its modules are executed where they lie.

Four third-party imports of the reference are absent from this image and are stubbed in
``sys.modules`` before ``network_models`` is imported:

* ``ptwt``                 -> ``oracle.haar.wavedec3 / waverec3`` (restatement of ptwt 0.1.9, parity unpinned there)
* ``timm.models.layers``   -> ``DropPath`` (per-sample stochastic depth, identity in eval), ``to_2tuple``,
                              ``trunc_normal_`` (== ``torch.nn.init.trunc_normal_``)
* ``torchinfo``, ``ptflops`` -> import-only no-ops (``network_backbone.py:22-23``, ``waveformer.py:14``)
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("WAVEFORMER_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network_models"))


class _DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _install_stubs() -> None:
    from . import haar

    if "ptwt" not in sys.modules:
        m = types.ModuleType("ptwt")
        m.wavedec3 = haar.wavedec3
        m.waverec3 = haar.waverec3
        sys.modules["ptwt"] = m
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models = models
        models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "torchinfo" not in sys.modules:
        m = types.ModuleType("torchinfo")
        m.summary = lambda *a, **k: None
        sys.modules["torchinfo"] = m
    if "ptflops" not in sys.modules:
        m = types.ModuleType("ptflops")
        m.get_model_complexity_info = lambda *a, **k: (None, None)
        sys.modules["ptflops"] = m


_CACHE = {}


def load_reference():
    """Return the reference ``network_models`` package (imported from REFERENCE_ROOT with the stubs in place)."""
    if "nm" in _CACHE:
        return _CACHE["nm"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib

        nm = importlib.import_module("network_models")
    if not os.path.abspath(nm.__file__).startswith(os.path.abspath(REFERENCE_ROOT)):
        raise RuntimeError(f"'network_models' resolved to {nm.__file__}, not the reference tree")
    _CACHE["nm"] = nm
    return nm


def load_reference_inferer():
    """Return the reference's vendored ``monai.inferers.SlidingWindowInferer`` class."""
    load_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from monai.inferers import SlidingWindowInferer  # vendored copy under REFERENCE_ROOT

    return SlidingWindowInferer


def default_model_kwargs(img: int = 128, in_chans: int = 4, out_chans: int = 4):
    """The BASELINE configuration."""
    return dict(
        img_size=(img, img, img),
        patch_size=2,
        in_chans=in_chans,
        out_chans=out_chans,
        depths=[2, 2, 2, 2],
        feat_size=[48, 96, 192, 384],
        num_heads=[3, 6, 12, 24],
        drop_path_rate=0.1,
    )
