"""Torch-level operators over the C ABI (include/waveformer_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every operator below launches hand-written sm_100a
kernels from ``waveformer_b200/lib/libwaveformer_b200.so`` on the CURRENT CUDA stream of the input's device and
raises if the tensor is not on a CUDA device (no CPU fallback).
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch

from . import _lib

DETAIL_KEYS = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")  # ptwt's key order (letter order = D, H, W)

LAUNCHES = 0  # number of kernel-launching C-ABI calls made by this process (bench.py reports it)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}      # wf_dtype: WF_F32, WF_BF16, WF_F16
HALF_TYPES = (torch.bfloat16, torch.float16)                          # the two 16-bit storage / tensor-core operand formats


def _dtype_code(t, allow_f16: bool = True) -> int:
    dt = t if isinstance(t, torch.dtype) else t.dtype
    code = _CODES.get(dt)
    if code is None or (code == 2 and not allow_f16):
        # same convention as ptwt, which raises ValueError for dtypes it does not support
        raise ValueError(f"waveformer_b200: dtype {dt} not supported (float32, bfloat16 or float16)")
    return code


def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("waveformer_b200 operators run on CUDA tensors only (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("waveformer_b200: tensors live on different devices")
    return dev


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ====================================================================================================== Haar ====
def _dwt_ncdhw_raw(x: torch.Tensor, need_hf: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    dev = _need_cuda(x)
    if x.dim() < 3:
        raise ValueError("expected at least 3 dims [..., D, H, W]")
    D, H, W = x.shape[-3:]
    if D % 2 or H % 2 or W % 2:
        raise ValueError(f"waveformer_b200 Haar transform needs even extents, got {(D, H, W)}")
    x = x.contiguous()
    lead = tuple(x.shape[:-3])
    n = 1
    for v in lead:
        n *= v
    ll = torch.empty(lead + (D // 2, H // 2, W // 2), dtype=x.dtype, device=dev)
    hf = torch.empty((7,) + tuple(ll.shape), dtype=x.dtype, device=dev) if need_hf else None
    with torch.cuda.device(dev):
        st = _lib.lib().wf_dwt3d_ncdhw(x.data_ptr(), ll.data_ptr(), _ptr(hf), _dtype_code(x), n, D, H, W, ll.numel(),
                                       _stream(dev))
    _lib.check(st, "wf_dwt3d_ncdhw")
    _count()
    return ll, hf


def _idwt_ncdhw_raw(ll: torch.Tensor, hf: Optional[torch.Tensor], gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _need_cuda(ll, hf, gate)
    ll = ll.contiguous()
    d, h, w = ll.shape[-3:]
    lead = tuple(ll.shape[:-3])
    n = 1
    for v in lead:
        n *= v
    if hf is not None:
        hf = hf.contiguous()
        if tuple(hf.shape) != (7,) + tuple(ll.shape) or hf.dtype != ll.dtype:
            raise ValueError("detail stack must be [7, *ll.shape] with ll's dtype")
    if gate is not None:
        gate = gate.contiguous()
        if hf is None or gate.shape != hf.shape or gate.dtype != ll.dtype:
            raise ValueError("gate must match the detail stack")
    x = torch.empty(lead + (2 * d, 2 * h, 2 * w), dtype=ll.dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_idwt3d_ncdhw(ll.data_ptr(), _ptr(hf), _ptr(gate), x.data_ptr(), _dtype_code(ll), n, d, h, w,
                                        ll.numel(), _stream(dev))
    _lib.check(st, "wf_idwt3d_ncdhw")
    _count()
    return x


def _dwt_ndhwc_raw(x: torch.Tensor, need_hf: bool, hf_dtype: Optional[torch.dtype] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    dev = _need_cuda(x)
    if x.dim() != 5:
        raise ValueError("expected [B, D, H, W, C]")
    B, D, H, W, C = x.shape
    if D % 2 or H % 2 or W % 2:
        raise ValueError(f"waveformer_b200 Haar transform needs even extents, got {(D, H, W)}")
    xs = _voxel_stride(x)
    if xs is None:
        x = x.contiguous()
        xs = C
    ll = torch.empty((B, D // 2, H // 2, W // 2, C), dtype=x.dtype, device=dev)
    hf = torch.empty((7,) + tuple(ll.shape), dtype=hf_dtype or x.dtype, device=dev) if need_hf else None
    with torch.cuda.device(dev):
        st = _lib.lib().wf_dwt3d_ndhwc(x.data_ptr(), ll.data_ptr(), _ptr(hf), _dtype_code(x),
                                       _dtype_code(hf) if need_hf else _dtype_code(x), B, D, H, W, C, xs, C, ll.numel(),
                                       _stream(dev))
    _lib.check(st, "wf_dwt3d_ndhwc")
    _count()
    return ll, hf


def _voxel_stride(t: torch.Tensor) -> Optional[int]:
    """Voxel stride of a [B, D, H, W, C] tensor that is dense over voxels with a (possibly wider) channel pitch."""
    B, D, H, W, C = t.shape
    s = t.stride()
    vs = s[3]
    if s[4] == 1 and vs >= C and s[2] == W * vs and s[1] == H * W * vs and (B == 1 or s[0] == D * H * W * vs):
        return vs
    return None


def _idwt_ndhwc_raw(ll: torch.Tensor, hf: Optional[torch.Tensor], gate: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _need_cuda(ll, hf, gate, out)
    if ll.dim() != 5:
        raise ValueError("expected ll [B, d, h, w, C]")
    B, d, h, w, C = ll.shape
    lls = _voxel_stride(ll)
    if lls is None:
        ll = ll.contiguous()
        lls = C
    if hf is not None:
        hf = hf.contiguous()
        if tuple(hf.shape) != (7, B, d, h, w, C) or hf.dtype != ll.dtype:
            raise ValueError("detail stack must be [7, B, d, h, w, C] with ll's dtype")
    if gate is not None:
        gate = gate.contiguous()
        if hf is None or gate.shape != hf.shape or gate.dtype != ll.dtype:
            raise ValueError("gate must match the detail stack")
    if out is None:
        out = torch.empty((B, 2 * d, 2 * h, 2 * w, C), dtype=ll.dtype, device=dev)
    if tuple(out.shape) != (B, 2 * d, 2 * h, 2 * w, C) or out.dtype != ll.dtype:
        raise ValueError("out must be [B, 2d, 2h, 2w, C] with ll's dtype")
    xs = _voxel_stride(out)
    if xs is None:
        raise ValueError("out must be voxel-dense with channel stride 1 (a channel slice of an NDHWC buffer)")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_idwt3d_ndhwc(ll.data_ptr(), _ptr(hf), _ptr(gate), out.data_ptr(), _dtype_code(ll), B, d, h, w,
                                        C, lls, B * d * h * w * C, xs, _stream(dev))
    _lib.check(st, "wf_idwt3d_ndhwc")
    _count()
    return out


class _DwtNCDHW(torch.autograd.Function):
    """Orthonormal transform: the adjoint of analysis is synthesis, so backward is one IDWT launch."""

    @staticmethod
    def forward(ctx, x, need_hf):
        ll, hf = _dwt_ncdhw_raw(x, need_hf)
        ctx.need_hf = need_hf
        if need_hf:
            return ll, hf
        return ll, x.new_empty(0)

    @staticmethod
    def backward(ctx, g_ll, g_hf):
        hf = g_hf if (ctx.need_hf and g_hf is not None and g_hf.numel()) else None
        return _idwt_ncdhw_raw(g_ll, hf), None


class _IdwtNCDHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ll, hf, gate):
        ctx.save_for_backward(hf if gate is not None else None, gate)
        return _idwt_ncdhw_raw(ll, hf, gate)

    @staticmethod
    def backward(ctx, g):
        hf, gate = ctx.saved_tensors
        g_ll, g_c = _dwt_ncdhw_raw(g, True)
        if gate is None:
            return g_ll, g_c, None
        return g_ll, g_c * gate, g_c * hf


class _DwtNDHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, need_hf, hf_dtype=None):
        ll, hf = _dwt_ndhwc_raw(x, need_hf, hf_dtype)
        ctx.need_hf = need_hf
        if need_hf:
            return ll, hf
        return ll, x.new_empty(0)

    @staticmethod
    def backward(ctx, g_ll, g_hf):
        hf = g_hf if (ctx.need_hf and g_hf is not None and g_hf.numel()) else None
        if hf is not None and hf.dtype != g_ll.dtype:
            hf = hf.to(g_ll.dtype)
        return _idwt_ndhwc_raw(g_ll, hf), None, None


class _IdwtNDHWC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ll, hf, gate, out):
        ctx.save_for_backward(hf if gate is not None else None, gate)
        res = _idwt_ndhwc_raw(ll, hf, gate, out)
        if out is not None:
            ctx.mark_dirty(out)
        return res

    @staticmethod
    def backward(ctx, g):
        hf, gate = ctx.saved_tensors
        g_ll, g_c = _dwt_ndhwc_raw(g, True)
        if gate is None:
            return g_ll, g_c, None, None
        return g_ll, g_c * gate, g_c * hf, None


def dwt3d(x: torch.Tensor, need_hf: bool = True):
    """One Haar level on ``x[..., D, H, W]`` -> ``(ll, hf)``; ``hf`` is ``[7, *ll.shape]`` in DETAIL_KEYS order."""
    ll, hf = _DwtNCDHW.apply(x, need_hf)
    return ll, (hf if need_hf else None)


def idwt3d(ll: torch.Tensor, hf: Optional[torch.Tensor], gate: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _IdwtNCDHW.apply(ll, hf, gate)


def dwt3d_channels_last(x: torch.Tensor, need_hf: bool = True, hf_dtype: Optional[torch.dtype] = None):
    """One Haar level on channels-last ``x[B, D, H, W, C]`` -> ``ll[B, d, h, w, C]``, ``hf[7, B, d, h, w, C]``
    (``hf_dtype``: store the details as bf16 while x / ll stay fp32)."""
    ll, hf = _DwtNDHWC.apply(x, need_hf, hf_dtype)
    return ll, (hf if need_hf else None)


def idwt3d_channels_last(ll: torch.Tensor, hf: Optional[torch.Tensor], gate: Optional[torch.Tensor] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Synthesis into ``out`` (may be a channel slice ``buf[..., :C]`` of a wider NDHWC buffer = fused concat)."""
    return _IdwtNDHWC.apply(ll, hf, gate, out)


# ================================================================================================= attention ====
def relpos_bias_expand(table: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """Dense transposed bias ``bias_t[h, j, i] = table[index[i, j], h]`` (fp32) for wf_window_attn_fwd."""
    dev = _need_cuda(table, index)
    if index.dtype != torch.int64 or index.dim() != 2 or index.shape[0] != index.shape[1]:
        raise ValueError("relative_position_index must be int64 [N, N]")
    table = table.contiguous()
    index = index.contiguous()
    heads, n = table.shape[1], index.shape[0]
    out = torch.empty((heads, n, n), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_relpos_bias_expand(table.data_ptr(), _dtype_code(table), index.data_ptr(), out.data_ptr(),
                                              heads, n, table.shape[0], _stream(dev))
    _lib.check(st, "wf_relpos_bias_expand")
    _count()
    return out


_ATTN_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def relpos_bias_image(table: torch.Tensor, index: torch.Tensor, fmt: torch.dtype) -> torch.Tensor:
    """Dense 16-bit bias image ``img[h, i, 520] = fmt(table[index[i, j], h] * log2 e)`` for the tensor-core kernels."""
    dev = _need_cuda(table, index)
    if index.dtype != torch.int64 or tuple(index.shape) != (512, 512):
        raise ValueError("the tensor-core attention path needs a [512, 512] int64 relative_position_index")
    if fmt not in (torch.bfloat16, torch.float16):
        raise ValueError("bias image format must be bfloat16 or float16")
    table = table.contiguous()
    index = index.contiguous()
    heads = table.shape[1]
    L = _lib.lib()
    img = torch.empty(L.wf_relpos_bias_image_bytes(heads, 512) // 2, dtype=fmt, device=dev)
    with torch.cuda.device(dev):
        st = L.wf_relpos_bias_image(table.data_ptr(), _dtype_code(table), index.data_ptr(), img.data_ptr(),
                                    _ATTN_CODE[fmt], heads, 512, table.shape[0], _stream(dev))
    _lib.check(st, "wf_relpos_bias_image")
    _count()
    return img


def window_attention_uses_tensor_cores(grid, C: int, heads: int, ws: int, compute_dtype: torch.dtype) -> bool:
    if compute_dtype not in (torch.bfloat16, torch.float16):
        return False
    return bool(_lib.lib().wf_window_attn_tc_supported(int(grid[0]), int(grid[1]), int(grid[2]), int(C), int(heads), int(ws)))


_SPLIT_CACHE = {}


def split_cached(p: torch.Tensor):
    """``p`` as an error-compensated fp16 pair ``(hi, lo)``, ``hi = fp16(p)``, ``lo = fp16(p - hi)`` (22 significant bits
    together); cached like ``cast_cached`` until ``p`` is modified, moved or freed."""
    def build():
        f = p.detach().float().contiguous()
        hi = f.half()
        return hi, (f - hi.float()).half()

    return _pack_cached(_SPLIT_CACHE, id(p), (p._version, p.data_ptr(), p.device, p.dtype, tuple(p.shape)), p, build)


def _window_attention_split(x, qkv_w, qkv_b, proj_w, proj_b, bias_img, heads: int, ws: int, scale: float) -> torch.Tensor:
    """wf_window_attn_fwd_split: compensated fp16 operands on the tensor cores, fp32 in / out."""
    dev = _need_cuda(x, qkv_w, qkv_b, proj_w, proj_b, bias_img)
    if x.dim() != 5:
        raise ValueError("expected x [B, D1, H1, W1, C]")
    if qkv_b is None or bias_img is None or bias_img.dtype != torch.float16:
        raise ValueError("the compensated attention path needs a qkv bias and the fp16 bias image")
    x = x.float().contiguous()
    B, D1, H1, W1, C = x.shape
    L = _lib.lib()
    nbytes = L.wf_window_attn_split_workspace_bytes(B, D1, H1, W1, C, heads, ws)
    if nbytes == 0:
        raise ValueError(f"window attention: unsupported geometry grid={(D1, H1, W1)} C={C} heads={heads} ws={ws}")
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(x.shape, dtype=torch.float32, device=dev)
    (qh, ql), (ph, pl) = split_cached(qkv_w), split_cached(proj_w)
    with torch.cuda.device(dev):
        st = L.wf_window_attn_fwd_split(x.data_ptr(), qh.data_ptr(), ql.data_ptr(), f32_cached(qkv_b).data_ptr(),
                                        ph.data_ptr(), pl.data_ptr(), f32_cached(proj_b).data_ptr(), bias_img.data_ptr(),
                                        out.data_ptr(), work.data_ptr(), nbytes, B, D1, H1, W1, C, heads, ws, float(scale),
                                        _stream(dev))
    _lib.check(st, "wf_window_attn_fwd_split")
    _count(3)
    return out


def _window_attention_raw(x, qkv_w, qkv_b, proj_w, proj_b, bias_t, heads: int, ws: int, scale: float,
                          compute_dtype: Optional[torch.dtype] = None, bias_img: Optional[torch.Tensor] = None,
                          out_dtype: Optional[torch.dtype] = None, return_workspace: bool = False):
    """``compute_dtype`` (default: x.dtype) is the weight / GEMM-operand type: bf16 or fp16 -> tcgen05 tensor-core path
    (needs ``bias_img``; x may be fp32 or bf16; the result is ``out_dtype`` = compute dtype or fp32); fp32 -> CUDA-core
    fp32 path (needs ``bias_t``)."""
    dev = _need_cuda(x, qkv_w, qkv_b, proj_w, proj_b, bias_t, bias_img)
    if x.dim() != 5:
        raise ValueError("expected x [B, D1, H1, W1, C]")
    B, D1, H1, W1, C = x.shape
    cdt = compute_dtype or x.dtype
    if cdt not in _ATTN_CODE:
        raise ValueError(f"waveformer_b200: dtype {cdt} not supported (float32, bfloat16 or float16 operands)")
    _dtype_code(x)
    x = x.contiguous()
    code = _ATTN_CODE[cdt]
    if code == 0 and x.dtype != torch.float32:
        x = x.float()
    odt = out_dtype or cdt
    x_code = _dtype_code(x)
    L = _lib.lib()
    nbytes = L.wf_window_attn_workspace_bytes(code, B, D1, H1, W1, C, heads, ws)
    if nbytes == 0:
        raise ValueError(f"window attention: unsupported geometry grid={(D1, H1, W1)} C={C} heads={heads} ws={ws}")
    work = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(x.shape, dtype=odt, device=dev)
    qkv_w, qkv_b, proj_w, proj_b = (cast_cached(t, cdt) for t in (qkv_w, qkv_b, proj_w, proj_b))
    with torch.cuda.device(dev):
        st = L.wf_window_attn_fwd(x.data_ptr(), x_code, qkv_w.data_ptr(), _ptr(qkv_b), proj_w.data_ptr(),
                                  proj_b.data_ptr(), _ptr(bias_t), _ptr(bias_img), out.data_ptr(), _ATTN_CODE[odt],
                                  work.data_ptr(), nbytes, code, B, D1, H1, W1, C, heads, ws, float(scale), _stream(dev))
    _lib.check(st, "wf_window_attn_fwd")
    _count(3)
    return (out, work) if return_workspace else out


def _window_partition(x: torch.Tensor, ws: int) -> torch.Tensor:
    b, d, h, w, c = x.shape
    x = x.reshape(b, d // ws, ws, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).reshape(-1, ws * ws * ws, c)


def _window_unpartition(rows: torch.Tensor, shape, ws: int) -> torch.Tensor:
    """Inverse of _window_partition: window-major rows ``[B * nW * ws^3, C]`` -> ``[B, D, H, W, C]``."""
    b, d, h, w, c = shape
    t = rows.reshape(b, d // ws, h // ws, w // ws, ws, ws, ws, c)
    return t.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(b, d, h, w, c)


def _window_attention_backward(g, x, qkv_w, qkv_b, proj_w, proj_b, table, index, heads: int, ws: int, scale: float):
    """Gradients of window_attention in fp32.  q, k, v and O are recomputed by the fp32 forward kernels (nothing but the
    layer input is kept between forward and backward), the attention core is differentiated by wf_window_attn_bwd, and the
    gradients of the two Linear layers are library GEMMs on its d_qkv / on grad_out."""
    dev = x.device
    f32 = lambda t: None if t is None else t.detach().float().contiguous()
    xf, wq, bq, wp, bp, tb = (f32(t) for t in (x, qkv_w, qkv_b, proj_w, proj_b, table))
    index = index.contiguous()
    B, D1, H1, W1, C = xf.shape
    n = ws * ws * ws
    windows = B * (D1 // ws) * (H1 // ws) * (W1 // ws)
    M = windows * n
    bias_t = relpos_bias_expand(tb, index)
    _, work = _window_attention_raw(xf, wq, bq, wp, bp, bias_t, heads, ws, scale, torch.float32, None, torch.float32,
                                    return_workspace=True)
    wsf = work[: 4 * M * C * 4].view(torch.float32)
    o = wsf[3 * M * C:].view(M, C)
    g2 = g.detach().float().contiguous().view(M, C)       # the output buffer is window-major (reshape-only reverse)
    d_proj_w = g2.t() @ o
    d_proj_b = g2.sum(0)
    d_o = g2 @ wp
    d_qkv = torch.empty((M, 3 * C), dtype=torch.float32, device=dev)
    d_table = torch.zeros_like(tb)
    L = _lib.lib()
    stats = torch.empty(L.wf_window_attn_bwd_stats_bytes(windows, n, heads) // 4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = L.wf_window_attn_bwd(wsf.data_ptr(), bias_t.data_ptr(), tb.data_ptr(), index.data_ptr(), d_o.data_ptr(),
                                  d_qkv.data_ptr(), d_table.data_ptr(), stats.data_ptr(), windows, n, C, heads,
                                  tb.shape[0], float(scale), _stream(dev))
    _lib.check(st, "wf_window_attn_bwd")
    _count(2)
    xw = _window_partition(xf, ws).reshape(M, C)
    d_qkv_w = d_qkv.t() @ xw
    d_qkv_b = d_qkv.sum(0) if qkv_b is not None else None
    d_x = _window_unpartition(d_qkv @ wq, xf.shape, ws)
    return d_x, d_qkv_w, d_qkv_b, d_proj_w, d_proj_b, d_table


class _WindowAttention(torch.autograd.Function):
    """Forward = the CUDA kernels (any operand format).  Backward (training, BASELINE config 5) = fp32 recompute of
    q / k / v / O with the forward kernels + wf_window_attn_bwd; see _window_attention_backward."""

    @staticmethod
    def forward(ctx, x, qkv_w, qkv_b, proj_w, proj_b, table, index, bias_t, heads, ws, scale, compute_dtype, bias_img,
                out_dtype, split=False):
        ctx.save_for_backward(x, qkv_w, qkv_b, proj_w, proj_b, table, index)
        ctx.cfg = (heads, ws, scale)
        if split:
            return _window_attention_split(x, qkv_w, qkv_b, proj_w, proj_b, bias_img, heads, ws, scale)
        return _window_attention_raw(x, qkv_w, qkv_b, proj_w, proj_b, bias_t, heads, ws, scale, compute_dtype, bias_img,
                                     out_dtype)

    @staticmethod
    def backward(ctx, g):
        x, qkv_w, qkv_b, proj_w, proj_b, table, index = ctx.saved_tensors
        heads, ws, scale = ctx.cfg
        srcs = (x, qkv_w, qkv_b, proj_w, proj_b, table)
        grads = _window_attention_backward(g, x, qkv_w, qkv_b, proj_w, proj_b, table, index, heads, ws, scale)
        out = tuple(None if (s_ is None or g_ is None) else g_.to(s_.dtype) for s_, g_ in zip(srcs, grads))
        return out + (None,) * 9


def window_attention(x, qkv_w, qkv_b, proj_w, proj_b, table, index, bias_t, heads: int, ws: int, scale: float,
                     compute_dtype: Optional[torch.dtype] = None, bias_img: Optional[torch.Tensor] = None,
                     out_dtype: Optional[torch.dtype] = None, split: bool = False):
    """Window partition + attention + reshape-only reverse on channels-last ``x[B, D1, H1, W1, C]``.

    Returns the window-major result buffer viewed as ``[B, D1, H1, W1, C]`` - exactly what the reference produces at
    ``wave_helper.py:497-499`` (it never applies the inverse permute)."""
    return _WindowAttention.apply(x, qkv_w, qkv_b, proj_w, proj_b, table, index, bias_t, heads, ws, scale, compute_dtype,
                                  bias_img, out_dtype, split)



# ================================================================================================ block glue ====
def dwconv3d_channels_last(x: torch.Tensor, w27: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Depthwise 3x3x3 conv (padding 1) on ``x[B, D, H, W, C]``; ``w27`` fp32 ``[27, C]``, ``bias`` fp32 ``[C]``."""
    dev = _need_cuda(x, w27, bias)
    if x.dim() != 5:
        raise ValueError("expected [B, D, H, W, C]")
    x = x.contiguous()
    B, D, H, W, C = x.shape
    if w27.dtype != torch.float32 or tuple(w27.shape) != (27, C) or not w27.is_contiguous():
        raise ValueError("w27 must be contiguous fp32 [27, C]")
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_dwconv3d_ndhwc(x.data_ptr(), w27.data_ptr(), _ptr(bias), y.data_ptr(), _dtype_code(x), B, D, H,
                                          W, C, _stream(dev))
    _lib.check(st, "wf_dwconv3d_ndhwc")
    _count()
    return y


def dwconv3d_channels_last_stats(x: torch.Tensor, w27: torch.Tensor, bias: Optional[torch.Tensor], eps: float):
    """Depthwise conv + the (mean, rstd) of its result per (sample, channel) - one kernel when the bf16 tile kernel covers
    the geometry, else the convolution followed by the statistics pass.  Returns ``(y, mean_rstd[B * C * 2])``."""
    dev = _need_cuda(x, w27, bias)
    if x.dim() != 5:
        raise ValueError("expected [B, D, H, W, C]")
    x = x.contiguous()
    B, D, H, W, C = x.shape
    if x.dtype in HALF_TYPES and C % 8 == 0 and W >= 8 and w27.dtype == torch.float32 and tuple(w27.shape) == (27, C) \
            and w27.is_contiguous():
        y = torch.empty_like(x)
        sums = torch.empty(B * C * 2, dtype=torch.float64, device=dev)
        mr = torch.empty(B * C * 2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.lib().wf_dwconv3d_ndhwc_stats(x.data_ptr(), w27.data_ptr(), _ptr(bias), y.data_ptr(), sums.data_ptr(),
                                                    mr.data_ptr(), float(eps), _dtype_code(x), B, D, H, W, C, _stream(dev))
        if st == 0:
            _count(2)
            return y, mr
        if st != -7:        # WF_ERR_UNSUPPORTED falls through to the two-kernel form
            _lib.check(st, "wf_dwconv3d_ndhwc_stats")
    y = dwconv3d_channels_last(x, w27, bias)
    return y, _instnorm_stats(y, C, eps)


def repack_depthwise_weight(weight: torch.Tensor) -> torch.Tensor:
    """[C, 1, 3, 3, 3] conv weight -> fp32 [27, C] (tap-major) for ``dwconv3d_channels_last``."""
    c = weight.shape[0]
    return weight.detach().reshape(c, 27).t().contiguous().float()


_ACT = {"none": 0, None: 0, "relu": 1, "leakyrelu": 2, "lrelu": 2}


def _ndhwc_view(x: torch.Tensor):
    """[B, C, D, H, W] -> ([B, D, H, W, C] view, voxel stride); copies only if the tensor is not channels-last-dense."""
    v = x.permute(0, 2, 3, 4, 1)
    vs = _voxel_stride(v)
    if vs is None:
        v = v.contiguous()
        vs = v.shape[-1]
    return v, vs


def _instnorm_stats(v: torch.Tensor, vs: int, eps: float) -> torch.Tensor:
    dev = v.device
    B, D, H, W, C = v.shape
    sums = torch.empty(B * C * 2, dtype=torch.float64, device=dev)
    mr = torch.empty(B * C * 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_instnorm_stats_ndhwc(v.data_ptr(), sums.data_ptr(), mr.data_ptr(), _dtype_code(v), B,
                                                D * H * W, C, vs, float(eps), _stream(dev))
    _lib.check(st, "wf_instnorm_stats_ndhwc")
    _count(2)
    return mr


def instance_norm_stats(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """(mean, rstd) per (sample, channel) of ``x[B, C, D, H, W]`` as a flat fp32 tensor [B * C * 2]."""
    _need_cuda(x)
    v, vs = _ndhwc_view(x)
    return _instnorm_stats(v, vs, eps)


def instance_norm_act(x: torch.Tensor, act: str = "none", slope: float = 0.01, res: Optional[torch.Tensor] = None,
                      res_norm: bool = False, eps: float = 1e-5, out: Optional[torch.Tensor] = None,
                      gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
                      stats: Optional[torch.Tensor] = None, res_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(InstanceNorm(x) + R)`` for ``x[B, C, D, H, W]`` (any strides; channels-last-3d is copy-free), where ``R`` is
    nothing, ``res`` or ``InstanceNorm(res)``.  Returns a [B, C, D, H, W] tensor with channels-last-3d strides; ``out``
    (optional) is a [B, D, H, W, C] channels-last destination, e.g. a channel slice of a concat buffer."""
    dev = _need_cuda(x, res, out)
    v, vs = _ndhwc_view(x)
    B, D, H, W, C = v.shape
    # stats / res_stats: (mean, rstd) pairs already produced by the kernel that wrote x / res (fused statistics)
    mr = stats if stats is not None else _instnorm_stats(v, vs, eps)
    rv, rs, rmr = None, C, None
    if res is not None:
        if res.shape != x.shape or res.dtype != x.dtype:
            raise ValueError("residual must match x")
        rv, rs = _ndhwc_view(res)
        if res_norm:
            rmr = res_stats if res_stats is not None else _instnorm_stats(rv, rs, eps)
    if out is None:
        out = torch.empty((B, D, H, W, C), dtype=x.dtype, device=dev)
    ys = _voxel_stride(out)
    if ys is None or tuple(out.shape) != (B, D, H, W, C) or not (
            out.dtype == x.dtype or (x.dtype in (torch.float32, torch.float16) and out.dtype == torch.bfloat16)
            or (x.dtype == torch.float32 and out.dtype == torch.float16)):
        raise ValueError("out must be a voxel-dense [B, D, H, W, C] tensor of x's dtype (or 16-bit for fp32 x, bf16 for fp16 x)")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_instnorm_apply_ndhwc(v.data_ptr(), mr.data_ptr(), _ptr(rv), _ptr(rmr), _ptr(gamma), _ptr(beta),
                                                out.data_ptr(), _ACT[act], float(slope), _dtype_code(x),
                                                _dtype_code(out), B, D * H * W, C, vs, rs, ys, _stream(dev))
    _lib.check(st, "wf_instnorm_apply_ndhwc")
    _count()
    return out.permute(0, 4, 1, 2, 3)


def shortcut4_stats(xin: torch.Tensor, w1x1: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """(mean, rstd) per (sample, channel) of ``conv1x1(xin)`` for a FOUR-channel ``xin`` (fp32 or the weight's 16-bit type, dense
    channels-last-3d) as a flat fp32 tensor [B * C * 2], from the input's first and second moments - the convolution is not run."""
    dev = _need_cuda(xin, w1x1)
    iv, ivs = _ndhwc_view(xin)
    B, D, H, W, cin = iv.shape
    C = w1x1.shape[0]
    if cin != 4 or ivs != 4 or w1x1.dtype not in HALF_TYPES or tuple(w1x1.shape) != (C, 4, 1, 1, 1) or iv.dtype not in (torch.float32, w1x1.dtype):
        raise ValueError("shortcut4_stats: dense 4-channel channels-last volume (fp32 or the weight's type) and a 16-bit [C, 4, 1, 1, 1] weight")
    w4 = f32_cached(w1x1).reshape(C, 4)
    sums = torch.empty(14 * B, dtype=torch.float64, device=dev)
    mr = torch.empty(2 * B * C, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_shortcut4_stats(iv.data_ptr(), _dtype_code(iv), _dtype_code(w1x1), w4.data_ptr(), sums.data_ptr(), mr.data_ptr(),
                                           float(eps), B, D * H * W, C, _stream(dev))
    _lib.check(st, "wf_shortcut4_stats")
    _count(2)
    return mr


def instance_norm_act_shortcut4(x: torch.Tensor, xin: torch.Tensor, w1x1: torch.Tensor, stats: torch.Tensor, res_stats: torch.Tensor,
                                act: str = "leakyrelu", slope: float = 0.01, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``act(InstanceNorm(x) + InstanceNorm(conv1x1(xin)))`` for a 16-bit ``x[B, C, D, H, W]`` and a FOUR-channel ``xin``
    (fp32 or x's dtype, dense channels-last-3d): the shortcut of the network's first residual block, recomputed per voxel from
    ``w1x1`` ([C, 4, 1, 1, 1], x's dtype) instead of being read back.  ``stats`` / ``res_stats``: the (mean, rstd) tensors of
    ``x`` and of the shortcut as ``conv3d_c4_in_stats(..., store_shortcut=False)`` returns them."""
    dev = _need_cuda(x, xin, w1x1, out)
    v, vs = _ndhwc_view(x)
    iv, ivs = _ndhwc_view(xin)
    B, D, H, W, C = v.shape
    if x.dtype not in HALF_TYPES or w1x1.dtype != x.dtype or tuple(w1x1.shape) != (C, 4, 1, 1, 1):
        raise ValueError("instance_norm_act_shortcut4: 16-bit x and a [C, 4, 1, 1, 1] weight of the same type")
    if tuple(iv.shape) != (B, D, H, W, 4) or ivs != 4 or iv.dtype not in (torch.float32, x.dtype):
        raise ValueError("instance_norm_act_shortcut4: xin must be a dense 4-channel channels-last volume (fp32 or x's dtype)")
    if out is None:
        out = torch.empty((B, D, H, W, C), dtype=x.dtype, device=dev)
    ys = _voxel_stride(out)
    if ys is None or tuple(out.shape) != (B, D, H, W, C) or out.dtype != x.dtype:
        raise ValueError("out must be a voxel-dense [B, D, H, W, C] tensor of x's dtype")
    w4 = f32_cached(w1x1).reshape(C, 4)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_instnorm_apply_shortcut4_ndhwc(v.data_ptr(), stats.data_ptr(), iv.data_ptr(), _dtype_code(iv), w4.data_ptr(),
                                                          res_stats.data_ptr(), out.data_ptr(), _ACT[act], float(slope),
                                                          _dtype_code(x), B, D * H * W, C, vs, ys, _stream(dev))
    _lib.check(st, "wf_instnorm_apply_shortcut4_ndhwc")
    _count()
    return out.permute(0, 4, 1, 2, 3)


def instance_norm_act_head(x: torch.Tensor, head_w: torch.Tensor, head_b: Optional[torch.Tensor], act: str = "none",
                           slope: float = 0.01, res: Optional[torch.Tensor] = None, res_norm: bool = False,
                           eps: float = 1e-5, stats: Optional[torch.Tensor] = None,
                           res_stats: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``conv1x1(act(InstanceNorm(x) + R))`` for ``x[B, C, D, H, W]``: the normalised activation is consumed in registers
    and only the ``K`` logits are stored.  ``head_w``: the [K, C, 1, 1, 1] weight.  Returns [B, K, D, H, W] with
    channels-last-3d strides, ``out_dtype`` (default fp32)."""
    dev = _need_cuda(x, res, head_w, head_b)
    v, vs = _ndhwc_view(x)
    B, D, H, W, C = v.shape
    K = head_w.shape[0]
    mr = stats if stats is not None else _instnorm_stats(v, vs, eps)
    rv, rs, rmr = None, C, None
    if res is not None:
        if res.shape != x.shape or res.dtype != x.dtype:
            raise ValueError("residual must match x")
        rv, rs = _ndhwc_view(res)
        if res_norm:
            rmr = res_stats if res_stats is not None else _instnorm_stats(rv, rs, eps)
    odt = out_dtype or torch.float32
    out = torch.empty((B, D, H, W, K), dtype=odt, device=dev)
    hw = f32_cached(head_w).reshape(K, C)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_instnorm_apply_head_ndhwc(v.data_ptr(), mr.data_ptr(), _ptr(rv), _ptr(rmr), hw.data_ptr(),
                                                     _ptr(f32_cached(head_b)), out.data_ptr(), _ACT[act], float(slope),
                                                     _dtype_code(x), _dtype_code(out), B, D * H * W, C, K, vs, rs,
                                                     _stream(dev))
    _lib.check(st, "wf_instnorm_apply_head_ndhwc")
    _count()
    return out.permute(0, 4, 1, 2, 3)


def _pack_cached(cache: dict, key, tag, owner: torch.Tensor, build):
    """Repacked-weight cache shared by the tensor-core kernels: an entry is valid while its weak reference still points
    at the very same parameter object and the parameter has not been modified (``tag``); dead entries are evicted
    whenever the cache has grown since the last sweep, so it never outlives the models it served."""
    hit = cache.get(key)
    if hit is None or hit[0]() is not owner or hit[1] != tag:
        if len(cache) >= 64:
            for k in [k for k, v in cache.items() if v[0]() is None]:
                del cache[k]
        hit = (weakref.ref(owner), tag, build())
        cache[key] = hit
    return hit[2]


_K3_PACK = {}


def conv3d_k3_c48(x: torch.Tensor, weight: torch.Tensor, in_stats: Optional[torch.Tensor] = None, slope: float = 0.01,
                  eps: float = 1e-5, out: Optional[torch.Tensor] = None, stage_clocks: Optional[torch.Tensor] = None):
    """3^3 convolution 48 -> 48 (padding 1, no bias) of ``x[B, 48, D, H, 128]`` (bf16 or fp16, channels-last-3d strides) on
    the tensor cores.  With ``in_stats`` (the (mean, rstd) tensor of x) the input is InstanceNorm'd + LeakyReLU'd while it is
    staged, i.e. the call computes ``conv(lrelu(IN(x)))``.  Returns ``(y, stats)``: y bf16 [B, 48, D, H, 128] channels-last
    and the (mean, rstd) statistics of y for the following ``instance_norm_act(stats=...)``.  ``stage_clocks`` (diagnostic,
    fp16 only): an int64 [148, 3, 8] tensor that receives every CTA's per-role phase clocks (``wf_conv3d_k3_c48_stage_clocks``)."""
    dev = _need_cuda(x, weight, in_stats, out)
    v, vs = _ndhwc_view(x)
    B, D, H, W, C = v.shape
    if v.dtype not in HALF_TYPES or C != 48 or W != 128 or tuple(weight.shape) != (48, 48, 3, 3, 3):
        raise ValueError("conv3d_k3_c48: bf16 / fp16 [B, 48, D, H, 128] input and a [48, 48, 3, 3, 3] weight")
    fmt = v.dtype

    def build():
        w = weight.detach().float().permute(2, 3, 4, 0, 1).reshape(3, 3, 3, 48, 3, 2, 8)   # [dz, dy, dx, n, ks, chunk, e]
        return w.permute(0, 2, 4, 5, 1, 3, 6).contiguous().to(fmt)                         # [dz, dx, ks, chunk, dy, n, e]

    pack = _pack_cached(_K3_PACK, (id(weight), fmt), (weight._version, weight.data_ptr(), weight.dtype), weight, build)
    if out is None:
        out = torch.empty((B, D, H, W, 48), dtype=fmt, device=dev)
    ys = _voxel_stride(out)
    if ys is None or tuple(out.shape) != (B, D, H, W, 48) or out.dtype != fmt:
        raise ValueError("out must be a voxel-dense [B, D, H, 128, 48] tensor of the input's 16-bit type")
    sums = torch.empty(2 * B * 48, dtype=torch.float64, device=dev)
    mr = torch.empty(2 * B * 48, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        if stage_clocks is not None:
            if stage_clocks.dtype != torch.int64 or stage_clocks.numel() < 148 * 8 * (1 if in_stats is not None else 3) or not stage_clocks.is_contiguous():
                raise ValueError("stage_clocks: contiguous int64 tensor of at least 148 * 8 elements")
            st = _lib.lib().wf_conv3d_k3_c48_stage_clocks(v.data_ptr(), _dtype_code(v), pack.data_ptr(), out.data_ptr(), sums.data_ptr(),
                                                          mr.data_ptr(), _ptr(in_stats), float(slope), float(eps), B, D, H, W, vs,
                                                          ys, stage_clocks.data_ptr(), _stream(dev))
        else:
            st = _lib.lib().wf_conv3d_k3_c48_in_stats(v.data_ptr(), _dtype_code(v), pack.data_ptr(), out.data_ptr(), sums.data_ptr(),
                                                      mr.data_ptr(), _ptr(in_stats), float(slope), float(eps), B, D, H, W, vs,
                                                      ys, _stream(dev))
    _lib.check(st, "wf_conv3d_k3_c48_in_stats")
    _count(2)
    return out.permute(0, 4, 1, 2, 3), mr


def conv3d_k3_c96_c48(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None):
    """3^3 convolution 96 -> 48 (padding 1, no bias) of ``x[B, 96, D, H, 128]`` (bf16 / fp16, channels-last-3d strides - typically a
    decoder's concatenation buffer) as two passes of the 48 -> 48 tensor-core kernel: ``y1 = conv(x[:, :48], w[:, :48])``, then
    ``y = conv(x[:, 48:], w[:, 48:]) + y1`` with y1 added to the fp32 accumulators before rounding.  Returns ``(y, stats)`` like
    :func:`conv3d_k3_c48`."""
    dev = _need_cuda(x, weight, out)
    v, vs = _ndhwc_view(x)
    B, D, H, W, C = v.shape
    if v.dtype not in HALF_TYPES or C != 96 or W != 128 or tuple(weight.shape) != (48, 96, 3, 3, 3):
        raise ValueError("conv3d_k3_c96_c48: bf16 / fp16 [B, 96, D, H, 128] input and a [48, 96, 3, 3, 3] weight")
    fmt = v.dtype

    def build():
        w = weight.detach().float().permute(2, 3, 4, 0, 1).reshape(3, 3, 3, 48, 2, 3, 2, 8)   # [dz, dy, dx, n, half, ks, chunk, e]
        return w.permute(4, 0, 2, 5, 6, 1, 3, 7).contiguous().to(fmt)                         # [half][dz, dx, ks, chunk, dy, n, e]

    pack = _pack_cached(_K3_PACK, (id(weight), fmt, "c96"), (weight._version, weight.data_ptr(), weight.dtype), weight, build)
    if out is None:
        out = torch.empty((B, D, H, W, 48), dtype=fmt, device=dev)
    ys = _voxel_stride(out)
    if ys is None or tuple(out.shape) != (B, D, H, W, 48) or out.dtype != fmt:
        raise ValueError("out must be a voxel-dense [B, D, H, 128, 48] tensor of the input's 16-bit type")
    part = torch.empty((B, D, H, W, 48), dtype=fmt, device=dev)
    sums = torch.empty(2 * B * 48, dtype=torch.float64, device=dev)
    mr = torch.empty(2 * B * 48, dtype=torch.float32, device=dev)
    esz = v.element_size()
    with torch.cuda.device(dev):
        fn = _lib.lib().wf_conv3d_k3_c48_add_stats
        st = fn(v.data_ptr(), _dtype_code(v), pack[0].data_ptr(), None, part.data_ptr(), sums.data_ptr(), mr.data_ptr(), float(eps),
                B, D, H, W, vs, 0, 48, _stream(dev))
        _lib.check(st, "wf_conv3d_k3_c48_add_stats")
        st = fn(v.data_ptr() + 48 * esz, _dtype_code(v), pack[1].data_ptr(), part.data_ptr(), out.data_ptr(), sums.data_ptr(),
                mr.data_ptr(), float(eps), B, D, H, W, vs, 48, ys, _stream(dev))
        _lib.check(st, "wf_conv3d_k3_c48_add_stats")
    _count(4)
    return out.permute(0, 4, 1, 2, 3), mr


_CT_PACK = {}


def conv_transpose3d_k2s2(x: torch.Tensor, weight: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ConvTranspose3d(k=2, s=2, bias=False) of channels-last bf16 / fp16 ``x[B, D, H, W, Cin]`` with ``weight[Cin, Cout, 2, 2, 2]``
    into ``out[B, 2D, 2H, 2W, Cout]`` (may be a channel slice of a wider channels-last buffer)."""
    dev = _need_cuda(x, weight, out)
    if x.dtype not in HALF_TYPES or x.dim() != 5:
        raise ValueError("conv_transpose3d_k2s2 takes channels-last bf16 / fp16 [B, D, H, W, Cin]")
    xs = _voxel_stride(x)
    if xs is None:
        x = x.contiguous()
        xs = x.shape[-1]
    B, D, H, W, Cin = x.shape
    if tuple(weight.shape[2:]) != (2, 2, 2) or weight.shape[0] != Cin:
        raise ValueError("weight must be [Cin, Cout, 2, 2, 2]")
    Cout = weight.shape[1]
    fmt = x.dtype
    pack = _pack_cached(_CT_PACK, (id(weight), fmt), (weight._version, weight.data_ptr(), weight.dtype), weight,
                        lambda: weight.detach().permute(2, 3, 4, 1, 0).reshape(8 * Cout, Cin).to(fmt).contiguous())
    if out is None:
        out = torch.empty((B, 2 * D, 2 * H, 2 * W, Cout), dtype=x.dtype, device=dev)
    ys = _voxel_stride(out)
    if ys is None or tuple(out.shape) != (B, 2 * D, 2 * H, 2 * W, Cout) or out.dtype != x.dtype:
        raise ValueError("out must be a voxel-dense [B, 2D, 2H, 2W, Cout] tensor of x's type")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_convtranspose3d_k2s2_ndhwc(x.data_ptr(), pack.data_ptr(), out.data_ptr(), _dtype_code(x), B, D, H, W, Cin,
                                                      Cout, xs, ys, _stream(dev))
    _lib.check(st, "wf_convtranspose3d_k2s2_ndhwc")
    _count()
    return out


_C4_PACK = {}


def _pack_c4_weights(w3x3: torch.Tensor, w1x1: Optional[torch.Tensor], fmt: torch.dtype) -> torch.Tensor:
    """[n0, 4, 3, 3, 3] (+ [n1, 4, 1, 1, 1]) -> ``fmt`` [n0 + n1, 112] for wf_conv3d_c4_in_stats; cached per weight version."""
    key = (id(w3x3), None if w1x1 is None else id(w1x1), fmt)
    tag = (w3x3._version, w3x3.data_ptr(), None if w1x1 is None else (w1x1._version, w1x1.data_ptr()))

    def build():
        n0 = w3x3.shape[0]
        n1 = 0 if w1x1 is None else w1x1.shape[0]
        pack = torch.zeros((n0 + n1, 112), dtype=torch.float32, device=w3x3.device)
        pack[:n0, :108] = w3x3.detach().float().permute(0, 2, 3, 4, 1).reshape(n0, 108)
        if n1:
            pack[n0:, 52:56] = w1x1.detach().float().reshape(n1, 4)
        return pack.to(fmt).contiguous()

    return _pack_cached(_C4_PACK, key, tag, w3x3, build)


def conv3d_c4_in_stats(x: torch.Tensor, w3x3: torch.Tensor, w1x1: Optional[torch.Tensor] = None, eps: float = 1e-5,
                       out_dtype: Optional[torch.dtype] = None, store_shortcut: bool = True):
    """3^3 conv (+ optional 1^3 conv) of a 4-channel volume with fused InstanceNorm statistics.  ``x``: [B, 4, D, H, W]
    with channels-last-3d strides (fp32, bf16 or fp16).  Returns ``(y0, stats0, y1, stats1)``; ``y*`` are [B, n, D, H, W]
    channels-last-3d in ``out_dtype`` (bf16 / fp16 = the tensor-core operand format; default: the weight's 16-bit type,
    else x's, else bf16), ``stats*`` the (mean, rstd) tensors ``instance_norm_act(stats=...)`` takes.  ``store_shortcut=False``:
    ``y1`` is not written (returned as None) - only its statistics, for ``instance_norm_act_shortcut4``."""
    dev = _need_cuda(x, w3x3, w1x1)
    v, vs = _ndhwc_view(x)
    B, D, H, W, C = v.shape
    if C != 4 or vs != 4:
        raise ValueError("conv3d_c4_in_stats needs a dense 4-channel channels-last volume")
    if tuple(w3x3.shape[1:]) != (4, 3, 3, 3) or (w1x1 is not None and tuple(w1x1.shape[1:]) != (4, 1, 1, 1)):
        raise ValueError("weights must be [n0, 4, 3, 3, 3] and [n1, 4, 1, 1, 1]")
    n0 = w3x3.shape[0]
    n1 = 0 if w1x1 is None else w1x1.shape[0]
    fmt = out_dtype or (w3x3.dtype if w3x3.dtype in HALF_TYPES else (v.dtype if v.dtype in HALF_TYPES else torch.bfloat16))
    if fmt not in HALF_TYPES or (v.dtype != torch.float32 and v.dtype != fmt):
        raise ValueError("conv3d_c4_in_stats: 16-bit result format; x must be fp32 or already in that format")
    pack = _pack_c4_weights(w3x3, w1x1, fmt)
    y0 = torch.empty((B, D, H, W, n0), dtype=fmt, device=dev)
    y1 = torch.empty((B, D, H, W, n1), dtype=fmt, device=dev) if (n1 and store_shortcut) else None
    sums = torch.empty(2 * B * (n0 + n1), dtype=torch.float64, device=dev)
    mr0 = torch.empty(2 * B * n0, dtype=torch.float32, device=dev)
    mr1 = torch.empty(2 * B * n1, dtype=torch.float32, device=dev) if n1 else None
    with torch.cuda.device(dev):
        st = _lib.lib().wf_conv3d_c4_in_stats(v.data_ptr(), _dtype_code(v), _dtype_code(fmt), pack.data_ptr(), y0.data_ptr(), n0, n0, _ptr(y1),
                                              n1, n1, sums.data_ptr(), sums.data_ptr() + 16 * B * n0, mr0.data_ptr(),
                                              _ptr(mr1), float(eps), B, D, H, W, _stream(dev))
    _lib.check(st, "wf_conv3d_c4_in_stats")
    _count(3 if n1 else 2)
    return (y0.permute(0, 4, 1, 2, 3), mr0, None if y1 is None else y1.permute(0, 4, 1, 2, 3), mr1)


_CAST_CACHE = {}


def cast_cached(p: Optional[torch.Tensor], dtype: torch.dtype) -> Optional[torch.Tensor]:
    """``p`` in ``dtype`` (contiguous, detached).  The converted copy is cached until ``p`` is modified, moved or freed:
    an entry is valid only while its weak reference still points at the very same tensor object (``id`` values and
    device addresses are both recycled once a tensor dies, so neither can identify it)."""
    if p is None:
        return None
    if p.dtype == dtype and p.is_contiguous():
        return p.detach()
    key = (id(p), dtype)
    tag = (p._version, p.data_ptr(), p.device, p.dtype, tuple(p.shape))
    hit = _CAST_CACHE.get(key)
    if hit is None or hit[0]() is not p or hit[1] != tag:
        if len(_CAST_CACHE) > 4096:
            for k in [k for k, v in _CAST_CACHE.items() if v[0]() is None]:
                del _CAST_CACHE[k]
        hit = (weakref.ref(p), tag, p.detach().to(dtype).contiguous())
        _CAST_CACHE[key] = hit
    return hit[2]


def f32_cached(p: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """fp32 contiguous view / copy of a (small) parameter, cached like ``cast_cached``."""
    return cast_cached(p, torch.float32)


def layer_norm_cl(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float,
                  gelu: bool = False, out_dtype: Optional[torch.dtype] = None, also_bf16: bool = False):
    """LayerNorm over the last dim of a channels-last tensor (any leading dims, last-dim stride 1), optional GELU;
    ``out_dtype`` may differ from the input's (fp32 stream -> 16-bit operand).  ``also_bf16`` (True = bf16, or a 16-bit
    dtype) returns ``(y, y16)``: the same result also rounded to 16 bits by the same pass (GEMM operand + fp32 copy for
    the residual)."""
    dev = _need_cuda(x, weight, bias)
    C = x.shape[-1]
    if x.stride(-1) != 1:
        x = x.contiguous()
    x2 = x.reshape(-1, C) if x.is_contiguous() else x.contiguous().reshape(-1, C)
    rows = x2.shape[0]
    out_dtype = out_dtype or x.dtype
    y = torch.empty(x.shape, dtype=out_dtype, device=dev)
    dt2 = torch.bfloat16 if also_bf16 is True else also_bf16
    y2 = torch.empty(x.shape, dtype=dt2, device=dev) if also_bf16 else None
    with torch.cuda.device(dev):
        st = _lib.lib().wf_layernorm_ndhwc(x2.data_ptr(), _ptr(f32_cached(weight)), _ptr(f32_cached(bias)), y.data_ptr(),
                                           _ptr(y2), _dtype_code(dt2) if also_bf16 else 1, _dtype_code(x2), _dtype_code(y),
                                           rows, C, x2.stride(0), C,
                                           float(eps), int(gelu), _stream(dev))
    _lib.check(st, "wf_layernorm_ndhwc")
    _count()
    return (y, y2) if also_bf16 else y


_PE_PACK = {}


def patch_embed_k2s2(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """``Conv3d(4 -> C, kernel = stride = 2)`` of ``x[B, 4, D, H, W]`` (channels-last-3d strides) as the channels-last fp32
    stream ``[B, D/2, H/2, W/2, C]`` in one exact-fp32 pass.  Returns None when the geometry is outside the kernel (the
    caller keeps the library convolution)."""
    if (x.dim() != 5 or x.shape[1] != 4 or tuple(weight.shape[1:]) != (4, 2, 2, 2) or weight.shape[0] % 12 or weight.shape[0] > 384
            or x.dtype not in _CODES or any(v % 2 for v in x.shape[2:]) or weight.dtype != torch.float32):
        return None
    dev = _need_cuda(x, weight, bias)
    v, vs = _ndhwc_view(x)
    if vs != 4:
        return None
    B, D, H, W, _ = v.shape
    C = weight.shape[0]
    pack = _pack_cached(_PE_PACK, id(weight), (weight._version, weight.data_ptr(), weight.dtype), weight,
                        lambda: weight.detach().float().permute(2, 3, 4, 1, 0).reshape(32, C).contiguous())
    y = torch.empty((B, D // 2, H // 2, W // 2, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_patch_embed_k2s2_c4(v.data_ptr(), _dtype_code(v), pack.data_ptr(), _ptr(f32_cached(bias)), y.data_ptr(),
                                               B, D, H, W, C, _stream(dev))
    _lib.check(st, "wf_patch_embed_k2s2_c4")
    _count()
    return y


def ffn_fused_supported(x: torch.Tensor, C_hid: int, fmt: Optional[torch.dtype]) -> bool:
    """wf_ffn_front / wf_ffn_back cover the fp32-stream, 16-bit-operand FFN of encoder stages 1 and 2 (C = 48 / 96, 4C hidden)."""
    return (x.is_cuda and x.dtype == torch.float32 and fmt in HALF_TYPES and x.shape[-1] in (48, 96)
            and C_hid == 4 * x.shape[-1] and x.is_contiguous())


def ffn_front(x: torch.Tensor, norm2: torch.nn.LayerNorm, pw_weight: torch.Tensor, pw_bias: Optional[torch.Tensor],
              ln: torch.nn.LayerNorm, fmt: torch.dtype) -> torch.Tensor:
    """``GELU(ln(pwconv(norm2(x))))`` for the fp32 stream ``x[..., C]`` -> ``[..., 4C]`` in ``fmt`` (one kernel)."""
    dev = _need_cuda(x, pw_weight, pw_bias)
    C = x.shape[-1]
    rows = x.numel() // C
    t1 = torch.empty(x.shape[:-1] + (4 * C,), dtype=fmt, device=dev)
    w = cast_cached(pw_weight, fmt).view(4 * C, C)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_ffn_front(x.data_ptr(), _ptr(f32_cached(norm2.weight)), _ptr(f32_cached(norm2.bias)), float(norm2.eps),
                                     w.data_ptr(), _ptr(f32_cached(pw_bias)), _ptr(f32_cached(ln.weight)),
                                     _ptr(f32_cached(ln.bias)), float(ln.eps), t1.data_ptr(), _dtype_code(fmt), rows, C,
                                     _stream(dev))
    _lib.check(st, "wf_ffn_front")
    _count()
    return t1


def ffn_back(t2: torch.Tensor, ln: torch.nn.LayerNorm, fc_weight: torch.Tensor, fc_bias: Optional[torch.Tensor],
             x: torch.Tensor, norm2: torch.nn.LayerNorm) -> torch.Tensor:
    """``x + norm2(x) + fc(GELU(ln(t2)))`` -> fp32 ``[..., C]`` (one kernel; ``t2[..., 4C]`` 16-bit, ``x`` the fp32 stream)."""
    dev = _need_cuda(t2, fc_weight, fc_bias, x)
    C = x.shape[-1]
    rows = x.numel() // C
    if not t2.is_contiguous() or t2.shape[-1] != 4 * C or t2.numel() != rows * 4 * C:
        raise ValueError("ffn_back: t2 must be the dense [..., 4C] companion of x")
    out = torch.empty_like(x)
    w = cast_cached(fc_weight, t2.dtype)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_ffn_back(t2.data_ptr(), _dtype_code(t2), _ptr(f32_cached(ln.weight)), _ptr(f32_cached(ln.bias)),
                                    float(ln.eps), w.data_ptr(), _ptr(f32_cached(fc_bias)), x.data_ptr(),
                                    _ptr(f32_cached(norm2.weight)), _ptr(f32_cached(norm2.bias)), float(norm2.eps),
                                    out.data_ptr(), rows, C, _stream(dev))
    _lib.check(st, "wf_ffn_back")
    _count()
    return out


def _row_stride(t: torch.Tensor) -> Optional[int]:
    """Row pitch (elements) of a [..., N] tensor whose rows are dense and equally spaced (e.g. a channel slice of a channels-last
    buffer); None if the leading dimensions do not collapse into one."""
    if t.dim() < 2 or t.stride(-1) != 1:
        return None
    pitch = t.stride(-2)
    for i in range(t.dim() - 3, -1, -1):
        if t.shape[i] != 1 and t.stride(i) != t.stride(i + 1) * t.shape[i + 1]:
            return None
    return int(pitch) if pitch >= t.shape[-1] else None


def pw_gelu_dual_supported(h: torch.Tensor, u: Optional[torch.Tensor], n_out: int, out: torch.Tensor) -> bool:
    """Shapes / types ``pw_gelu_dual`` is built for (the reference configuration's learnable_up3 / learnable_up4).  ``u=None``:
    the variant whose residual arrives as an fp32 addend."""
    k2 = 0 if u is None else u.shape[-1]
    return (h.is_cuda and h.dtype in HALF_TYPES and out.dtype == h.dtype and h.is_contiguous()
            and (u is None or (u.dtype == h.dtype and u.is_contiguous() and u.numel() // k2 == h.numel() // h.shape[-1]))
            and (h.shape[-1], k2, n_out) in ((192, 0, 48), (192, 96, 48), (192, 192, 48)) and out.shape[-1] == n_out
            and h.numel() // h.shape[-1] == out.numel() // n_out
            and _row_stride(out) is not None and _row_stride(out) % 8 == 0 and out.data_ptr() % 16 == 0)


def pw_gelu_dual(h: torch.Tensor, u: Optional[torch.Tensor], w1: torch.Tensor, b1: Optional[torch.Tensor], w2: Optional[torch.Tensor],
                 b2: Optional[torch.Tensor], out: torch.Tensor, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[..., :N] = GELU(h) @ w1.T + b1 + u @ w2.T + b2 + addend`` in one tensor-core kernel (``wf_pw_gelu_dual``): the tail of
    ``ProjectionUpsample``.  ``h[..., K1]``, ``u[..., K2]`` dense 16-bit (``u``, ``w2``, ``b2`` may be None); ``addend`` fp32
    dense ``[..., N]`` or None; ``out[..., N]`` row-dense, possibly a channel slice of a wider channels-last buffer; ``w1`` / ``w2`` any
    tensor viewable as [N, K1] / [N, K2]."""
    dev = _need_cuda(h, u, w1, b1, w2, b2, out, addend)
    n = out.shape[-1]
    if not pw_gelu_dual_supported(h, u, n, out):
        raise ValueError("pw_gelu_dual: unsupported shapes / types (see pw_gelu_dual_supported)")
    k1, k2 = h.shape[-1], 0 if u is None else u.shape[-1]
    rows = h.numel() // k1
    if addend is not None and (addend.dtype != torch.float32 or not addend.is_contiguous() or addend.numel() != rows * n):
        raise ValueError("pw_gelu_dual: addend must be a dense fp32 [..., N] tensor")
    w1c = cast_cached(w1, h.dtype).reshape(n, k1)
    w2c = None if u is None else cast_cached(w2, h.dtype).reshape(n, k2)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_pw_gelu_dual(h.data_ptr(), _ptr(u), _dtype_code(h), w1c.data_ptr(), _ptr(f32_cached(b1)), _ptr(w2c),
                                        _ptr(f32_cached(b2)) if u is not None else None, _ptr(addend), out.data_ptr(), rows, k1, k2, n,
                                        _row_stride(out), _stream(dev))
    _lib.check(st, "wf_pw_gelu_dual")
    _count()
    return out


def split_f16(x: torch.Tensor):
    """fp32 ``x`` (any dense layout) -> ``(hi, lo)`` fp16 tensors of x's shape and strides with ``hi + lo == x`` to 22 bits."""
    dev = _need_cuda(x)
    if x.dtype != torch.float32 or x.numel() % 4:
        raise ValueError("split_f16: fp32 tensor with a multiple of 4 elements")
    hi = torch.empty_like(x, dtype=torch.float16)
    if hi.stride() != x.stride():          # not a dense tensor: fall back to its contiguous form
        x = x.contiguous()
        hi = torch.empty_like(x, dtype=torch.float16)
    lo = torch.empty_like(hi)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_split_f16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream(dev))
    _lib.check(st, "wf_split_f16")
    _count()
    return hi, lo


def gelu_(x: torch.Tensor) -> torch.Tensor:
    """In-place exact (erf) GELU of a contiguous bf16 / fp32 tensor."""
    dev = _need_cuda(x)
    if not x.is_contiguous():
        raise ValueError("gelu_ works in place on a contiguous tensor")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_gelu_inplace(x.data_ptr(), _dtype_code(x), x.numel(), _stream(dev))
    _lib.check(st, "wf_gelu_inplace")
    _count()
    return x


def residual_sum(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``a + b + c + bias`` over the last (channel) dim in one pass; a, b fp32 contiguous, c fp32 or bf16."""
    dev = _need_cuda(a, b, c, bias)
    if a.dtype != torch.float32 or b.dtype != torch.float32 or a.shape != b.shape or a.shape != c.shape:
        raise ValueError("residual_sum: a, b fp32 and all three of one shape")
    a, b, c = a.contiguous(), b.contiguous(), c.contiguous()
    C = a.shape[-1]
    out = torch.empty_like(a)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_residual_sum(a.data_ptr(), b.data_ptr(), c.data_ptr(), _dtype_code(c), _ptr(f32_cached(bias)),
                                        out.data_ptr(), a.numel() // C, C, _stream(dev))
    _lib.check(st, "wf_residual_sum")
    _count()
    return out


def groupnorm_fold_linear(mean_rstd: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor],
                          weight: torch.Tensor, bias: Optional[torch.Tensor], batch: int, out_dtype: torch.dtype):
    """GroupNorm(groups = C) followed by a linear map ``weight[N, C]``, folded per sample: returns ``(W'[B, N, C], b'[B, N])``
    in ``out_dtype`` with ``W' x + b' == weight @ GroupNorm(x) + bias`` for the statistics ``mean_rstd[B * C * 2]``."""
    dev = _need_cuda(mean_rstd, gamma, beta, weight, bias)
    N, C = weight.shape
    weight = weight.contiguous()
    wf_ = torch.empty((batch, N, C), dtype=out_dtype, device=dev)
    bf_ = torch.empty((batch, N), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_groupnorm_fold_linear(mean_rstd.data_ptr(), _ptr(gamma), _ptr(beta), weight.data_ptr(), _ptr(bias),
                                                 wf_.data_ptr(), bf_.data_ptr(), _dtype_code(weight), _dtype_code(wf_),
                                                 batch, C, N, _stream(dev))
    _lib.check(st, "wf_groupnorm_fold_linear")
    _count()
    return wf_, bf_


def patch_merge_layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, octants,
                           out_dtype: torch.dtype) -> Optional[torch.Tensor]:
    """Octant gather + LayerNorm over the 8C concatenation in one kernel: ``x[B, D, H, W, C]`` fp32 ->
    ``[B, D/2, H/2, W/2, 8C]``.  Returns None when the geometry is outside the kernel (the caller keeps torch.cat + LN)."""
    if (x.dtype != torch.float32 or not x.is_contiguous() or x.dim() != 5 or out_dtype not in _CODES
            or gamma is None or beta is None or gamma.dtype != torch.float32 or beta.dtype != torch.float32):
        return None
    B, D, H, W, C = x.shape
    if (D | H | W) & 1 or C % 16 or C > 192 or (8 * C // 128) not in (1, 2, 3, 4, 6, 8, 12) or len(octants) != 8:
        return None
    dev = _need_cuda(x, gamma, beta)
    code = 0
    for s, (i, j, k) in enumerate(octants):
        code |= ((i << 2) | (j << 1) | k) << (3 * s)
    y = torch.empty((B, D // 2, H // 2, W // 2, 8 * C), dtype=out_dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_patch_merge_layernorm(x.data_ptr(), gamma.contiguous().data_ptr(), beta.contiguous().data_ptr(),
                                                 y.data_ptr(), _dtype_code(y), B, D, H, W, C, code, float(eps), _stream(dev))
    _lib.check(st, "wf_patch_merge_layernorm")
    _count()
    return y


def upsample_trilinear_add(srcs, size, base: Optional[torch.Tensor] = None, align_corners: bool = False,
                           out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``base + sum_s trilinear(srcs[s] -> size)`` for channels-last ``[B, d, h, w, C]`` sources (1..3 of them)."""
    import ctypes

    srcs = [s.contiguous() for s in srcs]
    dev = _need_cuda(*srcs, base)
    B, C = srcs[0].shape[0], srcs[0].shape[-1]
    D, H, W = size
    for s in srcs:
        if s.dim() != 5 or s.shape[0] != B or s.shape[-1] != C or s.dtype != srcs[0].dtype:
            raise ValueError("sources must be [B, d, h, w, C] with one dtype")
    io_dtype = out_dtype or (base.dtype if base is not None else srcs[0].dtype)
    bs = C
    if base is not None:
        if tuple(base.shape) != (B, D, H, W, C) or base.dtype != io_dtype:
            raise ValueError("base must be [B, D, H, W, C] of the output dtype")
        bs = _voxel_stride(base)
        if bs is None:
            base = base.contiguous()
            bs = C
    y = torch.empty((B, D, H, W, C), dtype=io_dtype, device=dev)
    n = len(srcs)
    ptrs = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
    dims = (ctypes.c_int * (3 * n))(*[v for s in srcs for v in s.shape[1:4]])
    with torch.cuda.device(dev):
        st = _lib.lib().wf_upsample_trilinear_add_ndhwc(ptrs, dims, n, _ptr(base), y.data_ptr(), _dtype_code(srcs[0]),
                                                        _dtype_code(y), int(align_corners), B, D, H, W, C, bs, C,
                                                        _stream(dev))
    _lib.check(st, "wf_upsample_trilinear_add_ndhwc")
    _count()
    return y

# ============================================================================================ sliding window ====
def _flip_mask(flip) -> int:
    """``flip``: 0..7 bit mask (bit 0 / 1 / 2 = mirror z / y / x) or an iterable of spatial axes (0, 1, 2)."""
    if flip is None:
        return 0
    if isinstance(flip, int):
        m = flip
    else:
        m = 0
        for a in flip:
            if a not in (0, 1, 2):
                raise ValueError(f"mirror axes are spatial axes 0, 1, 2; got {a}")
            m |= 1 << a
    if not 0 <= m <= 7:
        raise ValueError(f"flip mask must be in 0..7, got {m}")
    return m


def sw_gather(vol: torch.Tensor, starts: torch.Tensor, roi, dtype: torch.dtype, channels_last: bool, flip=0) -> torch.Tensor:
    dev = _need_cuda(vol, starts)
    if vol.dtype != torch.float32 or vol.dim() != 5:
        raise ValueError("volume must be fp32 [Bv, C, D, H, W]")
    vol = vol.contiguous()
    _, C, D, H, W = vol.shape
    n = starts.shape[0]
    r0, r1, r2 = roi
    shape = (n, r0, r1, r2, C) if channels_last else (n, C, r0, r1, r2)
    win = torch.empty(shape, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().wf_sw_gather(vol.data_ptr(), win.data_ptr(), starts.data_ptr(), n, _dtype_code(win),
                                     int(channels_last), C, D, H, W, r0, r1, r2, _flip_mask(flip), _stream(dev))
    _lib.check(st, "wf_sw_gather")
    _count()
    return win


def sw_accumulate(seg: torch.Tensor, acc: torch.Tensor, starts: torch.Tensor, gz, gy, gx, floor_w: float,
                  channels_last: bool, flip=0) -> None:
    dev = _need_cuda(seg, acc, starts, gz, gy, gx)
    seg = seg.contiguous()
    n = seg.shape[0]
    if channels_last:
        _, r0, r1, r2, K = seg.shape
    else:
        _, K, r0, r1, r2 = seg.shape
    _, K2, D, H, W = acc.shape
    if K2 != K or acc.dtype != torch.float32 or not acc.is_contiguous():
        raise ValueError("accumulator must be contiguous fp32 [Bv, K, D, H, W]")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_sw_accumulate(seg.data_ptr(), acc.data_ptr(), starts.data_ptr(), gz.data_ptr(), gy.data_ptr(),
                                         gx.data_ptr(), float(floor_w), n, _dtype_code(seg), int(channels_last), K, D,
                                         H, W, r0, r1, r2, _flip_mask(flip), _stream(dev))
    _lib.check(st, "wf_sw_accumulate")
    _count()


def sw_finalize(acc: torch.Tensor, all_starts: torch.Tensor, gz, gy, gx, floor_w: float, roi,
                labels: Optional[torch.Tensor] = None, z_range: Optional[Tuple[int, int]] = None, flip=0,
                dst: Optional[torch.Tensor] = None, dst_scale: float = 1.0, dst_add: bool = False) -> None:
    """Divide the accumulated volume(s) by the recomputed count map, in place; ``z_range`` limits it to planes [a, b).
    ``all_starts``: int32 [n, 4] window table (slot, z, y, x); slot -1 = the window exists in every volume of ``acc``.
    ``dst`` (optional, acc's shape): ``dst = (dst if dst_add else 0) + dst_scale * normalised`` (mirror-TTA mean)."""
    dev = _need_cuda(acc, all_starts, gz, gy, gx, labels, dst)
    Bv, K, D, H, W = acc.shape
    r0, r1, r2 = roi
    za, zb = (0, D) if z_range is None else (int(z_range[0]), int(z_range[1]))
    if not acc.is_contiguous() or acc.dtype != torch.float32:
        raise ValueError("accumulator must be contiguous fp32 [Bv, K, D, H, W]")
    if dst is not None and (dst.shape != acc.shape or dst.dtype != torch.float32 or not dst.is_contiguous()):
        raise ValueError("dst must be a contiguous fp32 tensor of the accumulator's shape")
    with torch.cuda.device(dev):
        st = _lib.lib().wf_sw_finalize(acc.data_ptr(), _ptr(labels), all_starts.data_ptr(), all_starts.shape[0],
                                       gz.data_ptr(), gy.data_ptr(), gx.data_ptr(), float(floor_w), Bv, K, D, H, W, r0,
                                       r1, r2, za, zb, _flip_mask(flip), _ptr(dst), float(dst_scale), int(bool(dst_add)),
                                       _stream(dev))
    _lib.check(st, "wf_sw_finalize")
    _count()
