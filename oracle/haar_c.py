"""Oracle (test infrastructure only): ctypes front-end for ``oracle/haar3d.c`` (plain-C closed-form Haar).

Used by the tests as a second, independent statement of the transform and by ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the CPU port timed on the host cores (the leading dimension is split over a thread
pool; ctypes drops the GIL during the call).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhaar3d_oracle.so")
_LIB: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "haar3d.c")):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        build()
        lib = ctypes.CDLL(_SO)
        for name in ("haar3d_dwt_f32", "haar3d_idwt_f32", "haar3d_dwt_f64", "haar3d_idwt_f64"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_long] * 4
        _LIB = lib
    return _LIB


def _suffix(a: np.ndarray) -> str:
    if a.dtype == np.float32:
        return "f32"
    if a.dtype == np.float64:
        return "f64"
    raise ValueError(f"dtype {a.dtype} not supported (float32/float64 only, as ptwt)")


def _run(fn, src: np.ndarray, dst: np.ndarray, n: int, dims, threads: int) -> None:
    if threads <= 1 or n < 2:
        fn(src.ctypes.data, dst.ctypes.data, n, *dims)
        return
    parts = np.array_split(np.arange(n), min(threads, n))
    s_stride, d_stride = src.strides[0], dst.strides[0]

    def work(idx):
        fn(src.ctypes.data + int(idx[0]) * s_stride, dst.ctypes.data + int(idx[0]) * d_stride, len(idx), *dims)

    with ThreadPoolExecutor(len(parts)) as pool:
        list(pool.map(work, [p for p in parts if len(p)]))


def dwt3d(x: np.ndarray, threads: int = 1) -> np.ndarray:
    """x[..., D, H, W] (even extents) -> [..., 8, D/2, H/2, W/2]; sub-band order aaa,aad,ada,add,daa,dad,dda,ddd."""
    x = np.ascontiguousarray(x)
    D, H, W = x.shape[-3:]
    if D % 2 or H % 2 or W % 2:
        raise ValueError("the C oracle handles even extents only")
    lead = x.shape[:-3]
    n = int(np.prod(lead)) if lead else 1
    src = x.reshape(n, D, H, W)
    out = np.empty((n, 8, D // 2, H // 2, W // 2), dtype=x.dtype)
    _run(getattr(_lib(), f"haar3d_dwt_{_suffix(x)}"), src, out, n, (D, H, W), threads)
    return out.reshape(*lead, 8, D // 2, H // 2, W // 2)


def idwt3d(c: np.ndarray, threads: int = 1) -> np.ndarray:
    """c[..., 8, d, h, w] -> x[..., 2d, 2h, 2w]."""
    c = np.ascontiguousarray(c)
    d, h, w = c.shape[-3:]
    lead = c.shape[:-4]
    n = int(np.prod(lead)) if lead else 1
    src = c.reshape(n, 8, d, h, w)
    out = np.empty((n, 2 * d, 2 * h, 2 * w), dtype=c.dtype)
    _run(getattr(_lib(), f"haar3d_idwt_{_suffix(c)}"), src, out, n, (d, h, w), threads)
    return out.reshape(*lead, 2 * d, 2 * h, 2 * w)
