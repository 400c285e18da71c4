"""Re-hosted sliding-window inferer (reference: the vendored MONAI ``SlidingWindowInferer`` /
``sliding_window_inference``, ``monai/inferers/inferer.py:382-535`` and ``monai/inferers/utils.py:43-321``, as
configured at ``4_predict.py:199-205``).

What changes relative to the reference's loop (same results):

* windows are gathered on the device by one kernel per batch (``wf_sw_gather``) straight into the layout and dtype the
  network consumes (channels-last bf16), instead of Python slicing + ``torch.cat``;
* the gaussian weighting and the scatter-add into the stitched volume are one kernel (``wf_sw_accumulate``), the
  accumulator stays in fp32 on the device;
* the count map is never materialised or reduced: it is geometry only, so ``wf_sw_finalize`` recomputes it per voxel;
* with a process group, the window list is SHARDED across ranks (one process per GPU) in contiguous balanced runs.
  A rank stitches its share into zero-initialised local volumes; a volume whose windows ended up on several ranks
  gets ONE ``reduce(SUM)`` (NCCL over NVLink) to its owner rank, which normalises it.  One volume on N GPUs is the
  north star's case (a single reduce of the stitched logit volume to rank 0); with at least N volumes no volume is
  split and the path has no collective at all.  Windows are independent, so nothing else is ever exchanged.
* host input / host output are STREAMED: a volume handed over in (pinned) host memory is copied in z-slabs on a copy
  stream and a window batch only waits for the slabs it reads; with ``device="cpu"`` (MONAI's name for "where the
  stitched output lives") a z-slab of the output is normalised and copied back as soon as no remaining window touches
  it, so most of both transfers hides behind the window forwards.
"""
from __future__ import annotations

import bisect
import itertools
import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops

__all__ = ["SlidingWindowInferer", "sliding_window_inference", "window_starts", "scan_interval", "gaussian_factors",
           "shard_batches", "volume_plan"]


# ----------------------------------------------------------------------------------------------- geometry (host)
def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: Sequence[float]) -> Tuple[int, ...]:
    """``_get_scan_interval`` (``monai/inferers/utils.py:363-384``)."""
    out = []
    for img, roi, o in zip(image_size, roi_size, overlap):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - o))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], interval: Sequence[int]) -> List[Tuple[int, ...]]:
    """Window origins in the order ``dense_patch_slices`` enumerates them (``monai/data/utils.py:171-211``)."""
    axes = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        if iv == 0:
            count = 1
        else:
            n = int(math.ceil(float(img) / iv))
            reach = [k for k in range(n) if k * iv + roi >= img]
            count = reach[0] + 1 if reach else 1
        axes.append([k * iv - max(k * iv + roi - img, 0) for k in range(count)])
    return list(itertools.product(*axes))


def gaussian_factors(roi_size: Sequence[int], mode: str, sigma_scale: Sequence[float]):
    """Per-axis factors of ``compute_importance_map`` (``monai/data/utils.py:1088-1138``) and the clamp floor
    ``max(min(map), 1e-3)``.  The kernels rebuild ``w = max((gz*gy)*gx, floor)`` from these three vectors."""
    mode = str(mode).lower()
    if mode.endswith("gaussian"):
        fac = []
        for n, s in zip(roi_size, sigma_scale):
            x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float)
            fac.append(torch.exp(x ** 2 / (-2 * (n * s) ** 2)))
    elif mode.endswith("constant"):
        fac = [torch.ones(n, dtype=torch.float) for n in roi_size]
    else:
        raise ValueError(f"Unsupported mode: {mode}, available options are ['constant', 'gaussian'].")
    corner = (fac[0].min() * fac[1].min()) * fac[2].min()  # same fp32 product order as the map itself
    floor = max(float(corner), 1e-3)
    return fac, floor


def shard_batches(num_windows: int, sw_batch_size: int, rank: int, world: int) -> List[range]:
    """Window batches owned by ``rank``.  The global window list (volume-major, as the reference enumerates it) is cut
    into consecutive ``sw_batch_size`` chunks and the chunks are split into ``world`` CONTIGUOUS, balanced runs, so a
    rank touches as few volumes as possible (one volume -> every rank shares it; >= world volumes -> none is shared)."""
    batches = [range(s, min(s + sw_batch_size, num_windows)) for s in range(0, num_windows, sw_batch_size)]
    nb = len(batches)
    return batches[rank * nb // world:(rank + 1) * nb // world]


def volume_plan(num_volumes: int, windows_per_volume: int, sw_batch_size: int, world: int):
    """For every volume: the sorted list of ranks that stitch at least one of its windows.  Pure geometry, identical
    on every rank.  ``owner(v) = ranks[0]`` finalises volume ``v``; volumes with more than one rank need the reduce."""
    total = num_volumes * windows_per_volume
    touch = [[] for _ in range(num_volumes)]
    for r in range(world):
        vols = sorted({i // windows_per_volume for b in shard_batches(total, sw_batch_size, r, world) for i in b})
        for v in vols:
            touch[v].append(r)
    return touch


# ------------------------------------------------------------------------------------------------------ inferer
def output_schedule(batches: Sequence[Sequence[Tuple[int, int]]], depth: int):
    """When is a z-slab of a volume final?  ``batches[j]`` lists ``(volume, z_start)`` of the windows of batch j, in
    processing order.  Returns ``sched[j] = [(volume, z_a, z_b), ...]``: after batch j no remaining window of `volume`
    starts below z_b, so planes [z_a, z_b) can be normalised and shipped.  The slabs of a volume tile [0, depth)."""
    remaining = {}
    for b in batches:
        for v, z in b:
            remaining.setdefault(v, []).append(z)
    done = {v: 0 for v in remaining}
    sched = []
    for b in batches:
        for v, z in b:
            remaining[v].remove(z)
        out = []
        for v in sorted({v for v, _ in b}):
            z_final = min(remaining[v]) if remaining[v] else depth
            if z_final > done[v]:
                out.append((v, done[v], z_final))
                done[v] = z_final
        sched.append(out)
    return sched


_COPY_STREAMS = {}
_PINNED = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def _pinned_out(shape, cache) -> torch.Tensor:
    """Page-locked result buffer, reused while the shape stays the same (pinning 143 MB costs more than the transfer):
    the tensor a host-output call returns is valid until the next call on the same inferer."""
    shape = tuple(shape)
    buf = cache.get("host_out")
    if buf is None or tuple(buf.shape) != shape:
        buf = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        cache["host_out"] = buf
    return buf


class SlidingWindowInferer:
    """Callable ``inferer(inputs, network)`` with MONAI's constructor arguments for the path the reference uses.

    Extra (keyword-only) arguments: ``process_group`` / ``shard`` (patch sharding over one process per GPU),
    ``compute_dtype`` (dtype the windows are handed to the network in; default: the network's parameter dtype),
    ``channels_last`` (hand the network channels-last-3d windows), ``return_labels`` (also produce the argmax map).

    ``device``: ``None`` keeps the stitched volume in HBM (a CUDA tensor is returned whatever the input's device);
    ``"cpu"`` returns it in page-locked host memory, streamed back slab by slab while later windows still run (the
    buffer is reused by the next call of this inferer).  ``sw_device`` is implied: windows always run on the GPU.
    """

    def __init__(self, roi_size, sw_batch_size: int = 1, overlap=0.25, mode="constant", sigma_scale=0.125,
                 padding_mode="constant", cval: float = 0.0, sw_device=None, device=None, progress: bool = False,
                 cache_roi_weight_map: bool = False, cpu_thresh=None, buffer_steps=None, buffer_dim: int = -1,
                 with_coord: bool = False, *, process_group=None, shard: bool = True, compute_dtype=None,
                 channels_last: bool = True, return_labels: bool = False):
        if buffer_steps:
            raise NotImplementedError("buffered stitching is a host-memory optimisation of the reference; the "
                                      "accumulator lives in HBM here")
        if with_coord:
            raise NotImplementedError("with_coord is not used on this path")
        self.roi_size = tuple(roi_size) if isinstance(roi_size, (tuple, list)) else (roi_size,) * 3
        self.sw_batch_size = int(sw_batch_size)
        self.overlap = overlap
        self.mode = getattr(mode, "value", mode)
        self.sigma_scale = sigma_scale
        self.padding_mode = getattr(padding_mode, "value", padding_mode)
        self.cval = cval
        self.process_group = process_group
        self.shard = shard
        self.compute_dtype = compute_dtype
        self.channels_last = channels_last
        self.return_labels = return_labels
        self.device = None if device is None else torch.device(device)
        if self.device is not None and self.device.type not in ("cpu", "cuda"):
            raise ValueError(f"device must be a cpu or cuda device, got {device}")
        self.labels: Optional[torch.Tensor] = None
        self.owned_volumes: List[int] = []
        self._geom_cache = {}

    def __call__(self, inputs: torch.Tensor, network: Callable[..., torch.Tensor], *args, **kwargs):
        out, labels, owned = _run(inputs, self.roi_size, self.sw_batch_size, network, self.overlap, self.mode,
                                  self.sigma_scale, self.padding_mode, self.cval, self.process_group, self.shard,
                                  self.compute_dtype, self.channels_last, self.return_labels, False, self._geom_cache,
                                  args, kwargs, host_out=self.device is not None and self.device.type == "cpu")
        self.labels = labels
        self.owned_volumes = owned      # indices (into the input batch) of the volumes returned on THIS rank
        return out


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, sw_device=None, device=None,
                             progress=False, roi_weight_map=None, process_fn=None, buffer_steps=None, buffer_dim=-1,
                             with_coord=False, *args, process_group=None, **kwargs):
    """Functional form with MONAI's signature (``monai/inferers/utils.py:43-64``)."""
    if roi_weight_map is not None or process_fn is not None or buffer_steps or with_coord:
        raise NotImplementedError("roi_weight_map / process_fn / buffer_steps / with_coord are not used on this path")
    roi = tuple(roi_size) if isinstance(roi_size, (tuple, list)) else (roi_size,) * 3
    out, _, _ = _run(inputs, roi, int(sw_batch_size), predictor, overlap, getattr(mode, "value", mode), sigma_scale,
                     getattr(padding_mode, "value", padding_mode), cval, process_group, True, None, True, False, False,
                     _PINNED, args, kwargs, host_out=device is not None and torch.device(device).type == "cpu")
    return out


def _network_dtype(network, fallback: torch.dtype) -> torch.dtype:
    params = getattr(network, "parameters", None)
    if params is not None:
        for p in params():
            if p.is_floating_point():
                return p.dtype
    return fallback


def _run(inputs, roi_size, sw_batch_size, network, overlap, mode, sigma_scale, padding_mode, cval, group, shard,
         compute_dtype, channels_last, return_labels, gather_result, cache, args, kwargs, host_out: bool = False):
    """Returns ``(logits, labels, owned)``: ``logits[i]`` is the stitched fp32 volume ``owned[i]`` (indices into the
    input batch).  Single process: ``owned`` is every volume, i.e. exactly the reference's return value."""
    if inputs.dim() != 5:
        raise ValueError("the B200 inferer handles 3D volumes: inputs must be [B, C, D, H, W]")
    nsp = 3
    ov = tuple(overlap) if isinstance(overlap, (tuple, list)) else (overlap,) * nsp
    for o in ov:
        if o < 0 or o >= 1:
            raise ValueError(f"overlap must be >= 0 and < 1, got {overlap}.")
    sg = tuple(sigma_scale) if isinstance(sigma_scale, (tuple, list)) else (sigma_scale,) * nsp
    world, rank = 1, 0
    dist = None
    if shard and (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = inputs.device if inputs.is_cuda else _default_device(network)
    batch = inputs.shape[0]
    orig = tuple(inputs.shape[2:])
    roi = tuple(int(r) if r and r > 0 else int(o) for r, o in zip(roi_size, orig))
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad = []
    for k in range(nsp - 1, -1, -1):
        diff = max(roi[k] - orig[k], 0)
        pad.extend([diff // 2, diff - diff // 2])

    key = (size, roi, ov, mode, sg, batch, str(dev), world, rank, sw_batch_size)
    if key not in cache:
        starts = window_starts(size, roi, scan_interval(size, roi, ov))
        nwin = len(starts)
        mine = shard_batches(batch * nwin, sw_batch_size, rank, world)
        touch = volume_plan(batch, nwin, sw_batch_size, world)
        local_vols = [v for v in range(batch) if rank in touch[v]]          # volumes this rank stitches into
        slot = {v: i for i, v in enumerate(local_vols)}
        host_tables = [[(slot[i // nwin],) + tuple(starts[i % nwin]) for i in b] for b in mine]
        my_tables = [torch.tensor(t, dtype=torch.int32, device=dev) for t in host_tables]
        # contiguous runs => at most the FIRST local volume is finalised by an earlier rank; the rest are ours
        owned = [v for v in local_vols if touch[v][0] == rank]
        own_off = len(local_vols) - len(owned)
        assert owned == local_vols[own_off:]
        fin_table = torch.tensor([(slot[v] - own_off,) + tuple(s) for v in owned for s in starts] or [(0, 0, 0, 0)],
                                 dtype=torch.int32, device=dev)
        shared = [v for v in range(batch) if len(touch[v]) > 1]
        fac, floor = gaussian_factors(roi, mode, sg)
        # streaming plans (host side): z-slabs of the input in the order windows need them, and when output slabs are final
        slab_ends = sorted({s_[0] + roi[0] for s_ in starts})
        need = [max((t[0], bisect.bisect_left(slab_ends, t[1] + roi[0])) for t in tb) for tb in host_tables]
        sched = output_schedule([[(t[0], t[1]) for t in tb] for tb in host_tables], size[0])
        fin_one = torch.tensor([(0,) + tuple(s_) for s_ in starts], dtype=torch.int32, device=dev)
        cache[key] = (my_tables, local_vols, slot, owned, own_off, fin_table, shared, touch, [f.to(dev) for f in fac], floor,
                      slab_ends, need, sched, fin_one)
    (my_tables, local_vols, slot, owned, own_off, fin_table, shared, touch, (gz, gy, gx), floor, slab_ends, need, sched,
     fin_one) = cache[key]

    # bring in only the volumes this rank touches (host input: this is the H2D copy of the end-to-end path)
    if local_vols:
        lo, hi = local_vols[0], local_vols[-1] + 1                          # contiguous by construction
        vol = inputs[lo:hi]
        in_events = None
        if not vol.is_cuda and vol.dtype == torch.float32 and not any(pad) and my_tables:
            # streamed H2D: z-slabs in the order the windows need them, one contiguous copy per (volume, channel, slab)
            src = vol.contiguous()
            vol = torch.empty(src.shape, dtype=torch.float32, device=dev)
            cs = _copy_stream(dev)
            cs.wait_stream(torch.cuda.current_stream(dev))
            in_events = {}
            with torch.cuda.stream(cs):
                for v in range(src.shape[0]):
                    a = 0
                    for si, b in enumerate(slab_ends):
                        for c in range(src.shape[1]):
                            vol[v, c, a:b].copy_(src[v, c, a:b], non_blocking=True)
                        in_events[(v, si)] = cs.record_event()
                        a = b
            vol.record_stream(cs)
        elif not vol.is_cuda:
            vol = vol.to(dev, non_blocking=True)
        if vol.dtype != torch.float32:
            vol = vol.float()
        if any(pad):
            vol = F.pad(vol, pad, mode=padding_mode, value=cval)
        vol = vol.contiguous()
    dtype = compute_dtype or _network_dtype(network, torch.float32)

    acc = None
    labels_buf = None
    # output slabs of volumes that are stitched by this rank alone can leave while later windows run
    stream_out = host_out and not any(pad)
    streamable = {slot[v] for v in owned if v not in shared} if stream_out else set()
    out_host, out_stream = None, None
    for j, st in enumerate(my_tables):
        if in_events is not None:
            torch.cuda.current_stream(dev).wait_event(in_events[need[j]])   # copies are in order: the last slab suffices
        win = ops.sw_gather(vol, st, roi, dtype, channels_last)
        if channels_last:
            win = win.permute(0, 4, 1, 2, 3)  # [n, C, r, r, r] with channels-last-3d strides, zero copy
        seg = network(win, *args, **kwargs)
        if not isinstance(seg, torch.Tensor):
            raise NotImplementedError("the B200 inferer stitches a single tensor output")
        if tuple(seg.shape[2:]) != roi:
            raise NotImplementedError("network output must have the window's spatial size")
        if acc is None:
            acc = torch.zeros((len(local_vols), seg.shape[1]) + size, dtype=torch.float32, device=dev)
        if seg.stride(1) == 1 and seg.shape[1] > 1:
            ops.sw_accumulate(seg.permute(0, 2, 3, 4, 1), acc, st, gz, gy, gx, floor, True)
        else:
            ops.sw_accumulate(seg, acc, st, gz, gy, gx, floor, False)
        for (sl, za, zb) in sched[j]:
            if sl not in streamable:
                continue
            if out_host is None:
                out_host = _pinned_out((len(owned), acc.shape[1]) + size, cache)
                out_stream = _copy_stream(dev)
                acc.record_stream(out_stream)
                if return_labels and labels_buf is None:
                    labels_buf = torch.empty((len(owned),) + size, dtype=torch.uint8, device=dev)
            i = sl - own_off
            ops.sw_finalize(acc[sl:sl + 1], fin_one, gz, gy, gx, floor, roi,
                            None if labels_buf is None else labels_buf[i:i + 1], z_range=(za, zb))
            out_stream.wait_event(torch.cuda.current_stream(dev).record_event())
            with torch.cuda.stream(out_stream):
                for k in range(acc.shape[1]):
                    out_host[i, k, za:zb].copy_(acc[sl, k, za:zb], non_blocking=True)

    if world > 1 and shared:
        k = _agree_channels(acc, network, dist, group, dev)
        if acc is None:
            acc = torch.zeros((0, k) + size, dtype=torch.float32, device=dev)
        zeros = None
        for v in shared:                              # one reduce(SUM) per volume that is split across ranks
            if v in slot:
                buf = acc[slot[v]]
            else:
                if zeros is None:
                    zeros = torch.empty((k,) + size, dtype=torch.float32, device=dev)
                buf = zeros.zero_()
            root = touch[v][0]
            dist.reduce(buf, dst=dist.get_global_rank(group, root) if group is not None else root,
                        op=dist.ReduceOp.SUM, group=group)
    labels = None
    if owned:
        acc_owned, fin = acc[own_off:], fin_table
        if return_labels:
            labels = labels_buf if labels_buf is not None else torch.empty((len(owned),) + size, dtype=torch.uint8, device=dev)
        rest = [v for v in owned if slot[v] not in streamable]        # everything, unless the output is streamed
        if len(rest) == len(owned):
            ops.sw_finalize(acc_owned, fin, gz, gy, gx, floor, roi, labels)
        else:
            for v in rest:
                i = slot[v] - own_off
                ops.sw_finalize(acc_owned[i:i + 1], fin_one, gz, gy, gx, floor, roi, None if labels is None else labels[i:i + 1])
        out = acc_owned
    else:
        out = None
    if out is not None and any(pad):
        crop = [slice(None), slice(None)]
        for sp in range(nsp):
            lo_ = pad[(nsp - 1 - sp) * 2]
            crop.append(slice(lo_, lo_ + orig[sp]))
        out = out[tuple(crop)]
        if labels is not None:
            labels = labels[tuple(crop[1:])]
    if host_out and out is not None:
        if out_host is None or any(pad):
            out_host = _pinned_out(out.shape, cache)
            out_host.copy_(out, non_blocking=True)
        else:
            for v in owned:                          # volumes that could not be streamed (finished by a reduce)
                if slot[v] not in streamable:
                    out_host[slot[v] - own_off].copy_(out[slot[v] - own_off], non_blocking=True)
            out_stream.synchronize()
        torch.cuda.current_stream(dev).synchronize()
        out = out_host
    return out, labels, owned


def _default_device(network) -> torch.device:
    params = getattr(network, "parameters", None)
    if params is not None:
        for p in params():
            return p.device
    return torch.device("cuda", torch.cuda.current_device())


def _agree_channels(acc, network, dist, group, dev) -> int:
    """Output channel count for ranks that stitched nothing (they still join the reduce with zeros)."""
    k = getattr(network, "out_chans", None)
    if k is not None:
        return int(k)
    t = torch.tensor([0 if acc is None else acc.shape[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())
