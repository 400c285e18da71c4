"""ctypes binding of the C ABI in include/waveformer_b200.h.

There is NO CPU fallback: if the CUDA library is missing or cannot be loaded, every op raises.  The library is only
*loaded* here (which works without a GPU, so the CPU test tier can check the exported symbols); launching anything
needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

from .build import LIB_PATH

_c = ctypes
_VOIDP, _I, _I64, _F, _SZ = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_size_t

# name -> (restype, argtypes); mirrors include/waveformer_b200.h one to one
SIGNATURES = {
    "wf_version": (_c.c_char_p, []),
    "wf_error_string": (_c.c_char_p, [_I]),
    "wf_last_cuda_error": (_I, []),
    "wf_dwt3d_ncdhw": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _I64, _I, _I, _I, _I64, _VOIDP]),
    "wf_dwt3d_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _I64, _VOIDP]),
    "wf_idwt3d_ncdhw": (_I, [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _I, _I64, _I, _I, _I, _I64, _VOIDP]),
    "wf_idwt3d_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _I64, _I64, _I64, _VOIDP]),
    "wf_relpos_bias_expand": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _I, _I, _I, _VOIDP]),
    "wf_window_attn_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "wf_relpos_bias_image_bytes": (_SZ, [_I, _I]),
    "wf_relpos_bias_image": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _I, _I, _I, _I, _VOIDP]),
    "wf_window_attn_tc_supported": (_I, [_I] * 6),
    "wf_window_attn_fwd": (_I, [_VOIDP, _I] + [_VOIDP] * 7 + [_I, _VOIDP, _SZ, _I, _I, _I, _I, _I, _I, _I, _I, _F, _VOIDP]),
    "wf_window_attn_split_workspace_bytes": (_SZ, [_I] * 7),
    "wf_window_attn_fwd_split": (_I, [_VOIDP] * 10 + [_SZ, _I, _I, _I, _I, _I, _I, _I, _F, _VOIDP]),
    "wf_window_attn_bwd_stats_bytes": (_SZ, [_I64, _I, _I]),
    "wf_window_attn_bwd": (_I, [_VOIDP] * 8 + [_I64, _I, _I, _I, _I, _F, _VOIDP]),
    "wf_patch_merge_layernorm": (_I, [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _c.c_uint32, _F, _VOIDP]),
    "wf_dwconv3d_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _VOIDP]),
    "wf_dwconv3d_ndhwc_stats": (_I, [_VOIDP] * 6 + [_F, _I, _I, _I, _I, _I, _I, _VOIDP]),
    "wf_instnorm_stats_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _I, _I64, _I, _I64, _F, _VOIDP]),
    "wf_instnorm_apply_ndhwc": (_I, [_VOIDP] * 7 + [_I, _F, _I, _I, _I, _I64, _I, _I64, _I64, _I64, _VOIDP]),
    "wf_instnorm_apply_shortcut4_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _I, _F, _I, _I, _I64, _I, _I64, _I64, _VOIDP]),
    "wf_shortcut4_stats": (_I, [_VOIDP, _I, _I, _VOIDP, _VOIDP, _VOIDP, _F, _I, _I64, _I, _VOIDP]),
    "wf_instnorm_apply_head_ndhwc": (_I, [_VOIDP] * 7 + [_I, _F, _I, _I, _I, _I64, _I, _I, _I64, _I64, _VOIDP]),
    "wf_groupnorm_fold_linear": (_I, [_VOIDP] * 7 + [_I, _I, _I, _I, _I, _VOIDP]),
    "wf_split_f16": (_I, [_VOIDP, _VOIDP, _VOIDP, _I64, _VOIDP]),
    "wf_gelu_inplace": (_I, [_VOIDP, _I, _I64, _VOIDP]),
    "wf_residual_sum": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _VOIDP, _VOIDP, _I64, _I, _VOIDP]),
    "wf_layernorm_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I64, _I, _I64, _I64, _F, _I, _VOIDP]),
    "wf_patch_embed_k2s2_c4": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _VOIDP]),
    "wf_ffn_front": (_I, [_VOIDP, _VOIDP, _VOIDP, _F, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _VOIDP, _I, _I64, _I, _VOIDP]),
    "wf_pw_gelu_dual": (_I, [_VOIDP, _VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _I64, _I, _I, _I, _I64, _VOIDP]),
    "wf_ffn_back": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _F, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _VOIDP, _I64, _I, _VOIDP]),
    "wf_upsample_trilinear_add_ndhwc": (_I, [_VOIDP, _VOIDP, _I, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _VOIDP]),
    "wf_conv3d_c4_in_stats": (_I, [_VOIDP, _I, _I, _VOIDP, _VOIDP, _I64, _I, _VOIDP, _I64, _I, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _I, _I, _I, _I, _VOIDP]),
    "wf_conv3d_k3_c48_in_stats": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _F, _I, _I, _I, _I, _I64, _I64, _VOIDP]),
    "wf_conv3d_k3_c48_add_stats": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _I, _I, _I, _I, _I64, _I64, _I64, _VOIDP]),
    "wf_conv3d_k3_c48_stage_clocks": (_I, [_VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _VOIDP, _F, _F, _I, _I, _I, _I, _I64, _I64, _VOIDP, _VOIDP]),
    "wf_convtranspose3d_k2s2_ndhwc": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _VOIDP]),
    "wf_sw_gather": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VOIDP]),
    "wf_sw_accumulate": (_I, [_VOIDP] * 6 + [_F, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VOIDP]),
    "wf_sw_finalize": (_I, [_VOIDP, _VOIDP, _VOIDP, _I, _VOIDP, _VOIDP, _VOIDP, _F] + [_I] * 11 + [_VOIDP, _F, _I, _VOIDP]),
}

_LIB: Optional[ctypes.CDLL] = None


class WaveformerB200Error(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (once) and return the CUDA library; raises if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise WaveformerB200Error(
                f"{LIB_PATH} is missing - build it with `python -m waveformer_b200.build` "
                "(waveformer_b200 has no CPU fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().wf_error_string(status).decode()
        if status == -2:
            raise ValueError(f"{what}: {msg}")
        raise WaveformerB200Error(f"{what}: {msg} (status {status})")
