"""CUDA-graph replay of the window forward.

One batch-2 window forward is ~450 kernel launches of 5-900 us; its CPU-side launch cost (12.6 ms) is as large as its GPU
time (12.7 ms), so the sliding-window loop is one host hiccup away from being launch-bound.  ``GraphedForward`` captures
``module(x)`` once per input signature (shape, dtype, strides) into a ``torch.cuda.CUDAGraph`` with static input / output
buffers and replays it afterwards: the 18 windows of a volume cost 9 replays + 9 gathers + 9 accumulates.

Everything the forward launches - cuDNN / cuBLAS calls and this package's C-ABI kernels, which take the CURRENT torch
stream - is capturable: no kernel synchronises, allocations go through torch's caching allocator (private graph pool),
the one-time ``cudaFuncSetAttribute`` opt-ins happen during the eager warm-up runs that precede the capture.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops

__all__ = ["GraphedForward"]


class GraphedForward:
    """Callable wrapper: ``y = GraphedForward(model)(x)``.  The returned tensor is a STATIC buffer that the next call
    overwrites - consume it (the inferer's accumulate kernel does) before calling again.  Inference only."""

    def __init__(self, module: nn.Module, warmup: int = 2):
        self.module = module
        self.warmup = int(warmup)
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor, torch.Tensor, int]] = {}
        self.out_chans = getattr(module, "out_chans", None)

    def parameters(self):
        return self.module.parameters()

    def eval(self):
        self.module.eval()
        return self

    def _key(self, x: torch.Tensor) -> Tuple:
        return (tuple(x.shape), x.dtype, tuple(x.stride()), x.device.index)

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("GraphedForward replays CUDA graphs: the input must be a CUDA tensor")
        key = self._key(x)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device)
            static_in.copy_(x)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                for _ in range(self.warmup):          # eager runs: cuDNN algorithm selection, lazy attribute opt-ins
                    self.module(static_in)
            torch.cuda.current_stream(x.device).wait_stream(side)
            torch.cuda.synchronize(x.device)
            graph = torch.cuda.CUDAGraph()
            before = ops.LAUNCHES
            with torch.cuda.graph(graph):
                static_out = self.module(static_in)
            entry = (graph, static_in, static_out, ops.LAUNCHES - before)   # own kernel launches recorded in the graph
            self._graphs[key] = entry
        graph, static_in, static_out, own_launches = entry
        if static_in.data_ptr() != x.data_ptr():
            static_in.copy_(x)
        graph.replay()
        ops._count(own_launches)          # keep the package's launch counter truthful under replay
        return static_out
