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
           "shard_windows", "shard_batches", "volume_plan", "volume_owner"]


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


def shard_windows(num_windows: int, rank: int, world: int, mode: str = "contiguous",
                  windows_per_volume: Optional[int] = None) -> List[int]:
    """Global window indices (volume-major, as the reference enumerates them) owned by ``rank``.

    ``"contiguous"``: ``world`` balanced consecutive runs cut at WINDOW granularity (sizes differ by at most one), so a
    rank touches as few volumes as possible - with ``volumes % world == 0`` no volume is split and the path has no
    collective; one volume on 8 ranks gives 2 or 3 windows each (18 = 6 x 2 + 2 x 3).
    ``"interleaved"``: window ``g`` goes to rank ``g % world`` - EVERY volume is split over all ranks (the north star's
    patch sharding: each volume costs one reduce of its stitched logits, issued while the next volume's windows run)."""
    if mode == "contiguous":
        return list(range(rank * num_windows // world, (rank + 1) * num_windows // world))
    if mode == "interleaved":
        return list(range(rank, num_windows, world))
    raise ValueError(f"shard mode must be 'contiguous' or 'interleaved', got {mode!r}")


def shard_batches(num_windows: int, sw_batch_size: int, rank: int, world: int, mode: str = "contiguous") -> List[List[int]]:
    """Window batches of ``rank``: its ``shard_windows`` share in ascending order, cut into ``sw_batch_size`` chunks."""
    mine = shard_windows(num_windows, rank, world, mode)
    return [mine[s:s + sw_batch_size] for s in range(0, len(mine), sw_batch_size)]


def volume_plan(num_volumes: int, windows_per_volume: int, sw_batch_size: int, world: int, mode: str = "contiguous"):
    """For every volume: the sorted list of ranks that stitch at least one of its windows.  Pure geometry, identical
    on every rank.  Volumes with more than one rank need the reduce; ``volume_owner`` says who receives it."""
    total = num_volumes * windows_per_volume
    touch = [[] for _ in range(num_volumes)]
    for r in range(world):
        for v in sorted({i // windows_per_volume for i in shard_windows(total, r, world, mode)}):
            touch[v].append(r)
    return touch


def volume_owner(v: int, touch: Sequence[Sequence[int]], mode: str = "contiguous") -> int:
    """Rank that finalises volume ``v``: the first rank that stitches it (contiguous runs), or - when every volume is
    split over all ranks - round-robin over the ranks that hold a share, so the outputs spread evenly."""
    ranks = touch[v]
    return ranks[0] if mode == "contiguous" else ranks[v % len(ranks)]


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
_GEOMETRY = {}          # plan cache of the functional API (bounded, geometry only - never an output buffer)
_PLAN_CACHE_ENTRIES = 8


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def _host_result(shape, holder: Optional[dict]) -> torch.Tensor:
    """Page-locked result buffer.  Default: a FRESH tensor per call (what MONAI returns).  ``holder`` (the inferer's
    ``reuse_output=True`` opt-in) keeps one buffer per shape alive across calls - pinning 143 MB costs more than moving
    it - at the price that the next call overwrites the tensor the previous one returned."""
    shape = tuple(shape)
    if holder is None:
        return torch.empty(shape, dtype=torch.float32, pin_memory=True)
    buf = holder.get("host_out")
    if buf is None or tuple(buf.shape) != shape:
        buf = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        holder["host_out"] = buf
    return buf


class _Plan:
    """Everything about one (geometry, sharding) combination that does not depend on the data: built once, cached."""

    def __init__(self, size, roi, ov, mode, sg, batch, dev, world, rank, sw_batch_size, shard_mode):
        starts = window_starts(size, roi, scan_interval(size, roi, ov))
        nwin = len(starts)
        self.starts, self.nwin = starts, nwin
        total = batch * nwin
        mine = shard_batches(total, sw_batch_size, rank, world, shard_mode)
        self.touch = touch = volume_plan(batch, nwin, sw_batch_size, world, shard_mode)
        self.owner = [volume_owner(v, touch, shard_mode) for v in range(batch)]
        local = [v for v in range(batch) if rank in touch[v]]               # volumes this rank stitches into
        self.owned = [v for v in local if self.owner[v] == rank]            # ... and returns
        # accumulator slots: owned volumes first, so the result is a leading view of the accumulator
        self.local_vols = self.owned + [v for v in local if self.owner[v] != rank]
        self.slot = slot = {v: i for i, v in enumerate(self.local_vols)}
        self.lo, self.hi = (min(local), max(local) + 1) if local else (0, 0)      # input volumes this rank reads
        self.host_tables = [[(slot[i // nwin],) + tuple(starts[i % nwin]) for i in b] for b in mine]
        self.batch_vols = [sorted({i // nwin for i in b}) for b in mine]
        self.tables = [torch.tensor(t, dtype=torch.int32, device=dev) for t in self.host_tables]
        # gather reads the input block [lo, hi): its own slot numbering
        self.in_tables = [torch.tensor([(i // nwin - self.lo,) + tuple(starts[i % nwin]) for i in b], dtype=torch.int32,
                                       device=dev) for b in mine]
        self.shared = [v for v in range(batch) if len(touch[v]) > 1]        # one reduce each, issued in this order
        # after which of MY batches is my share of volume v complete?  (-1: I never touch it)
        last = {}
        for j, vs in enumerate(self.batch_vols):
            for v in vs:
                last[v] = j
        self.done_after = [last.get(v, -1) for v in range(batch)]
        fac, self.floor = gaussian_factors(roi, mode, sg)
        self.fac = [f.to(dev) for f in fac]
        # streaming plans (host side): z-slabs of the input in the order windows need them, when output slabs are final
        self.slab_ends = sorted({s_[0] + roi[0] for s_ in starts})
        self.need = [max((i // nwin - self.lo, bisect.bisect_left(self.slab_ends, starts[i % nwin][0] + roi[0])) for i in b)
                     for b in mine]
        self.sched = output_schedule([[(t[0], t[1]) for t in tb] for tb in self.host_tables], size[0])
        # one volume's window list with slot -1 = "every volume of the call" (they share their geometry)
        self.fin = torch.tensor([(-1,) + tuple(s_) for s_ in starts], dtype=torch.int32, device=dev)


def _plan_for(cache: dict, key, build) -> "_Plan":
    plan = cache.pop(key, None)
    if plan is None:
        plan = build()
    cache[key] = plan                                   # most recently used last
    while len(cache) > _PLAN_CACHE_ENTRIES:
        cache.pop(next(iter(cache)))
    return plan


class SlidingWindowInferer:
    """Callable ``inferer(inputs, network)`` with MONAI's constructor arguments for the path the reference uses.

    Extra (keyword-only) arguments: ``process_group`` / ``shard`` (patch sharding over one process per GPU; ``shard`` may
    be ``True`` = ``"contiguous"``, ``"interleaved"`` or ``False``, see ``shard_windows``), ``compute_dtype`` (dtype the
    windows are handed to the network in; default: the network's parameter dtype), ``channels_last`` (hand the network
    channels-last-3d windows), ``return_labels`` (also produce the argmax map), ``reuse_output``.

    ``device``: ``None`` keeps the stitched volume in HBM (a CUDA tensor is returned whatever the input's device);
    ``"cpu"`` returns it in page-locked host memory, streamed back slab by slab while later windows still run - a fresh
    tensor per call like MONAI, or, with ``reuse_output=True``, one buffer that the next call of this inferer
    overwrites.  ``sw_device`` is implied: windows always run on the GPU.

    ``inferer(inputs, network, flip=axes)`` runs the pass on the volume mirrored along the spatial ``axes`` and returns
    the result mirrored back (one pass of the reference's test-time augmentation, ``light_training/prediction.py:
    129-156``) without building either mirrored copy; ``into=(dst, scale, add)`` folds ``dst (+)= scale * result`` into the
    normalisation kernel.
    """

    def __init__(self, roi_size, sw_batch_size: int = 1, overlap=0.25, mode="constant", sigma_scale=0.125,
                 padding_mode="constant", cval: float = 0.0, sw_device=None, device=None, progress: bool = False,
                 cache_roi_weight_map: bool = False, cpu_thresh=None, buffer_steps=None, buffer_dim: int = -1,
                 with_coord: bool = False, *, process_group=None, shard=True, compute_dtype=None,
                 channels_last: bool = True, return_labels: bool = False, reuse_output: bool = False):
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
        self._host_holder = {} if reuse_output else None

    def __call__(self, inputs: torch.Tensor, network: Callable[..., torch.Tensor], *args, flip=0, into=None, **kwargs):
        out, labels, owned = _run(inputs, self.roi_size, self.sw_batch_size, network, self.overlap, self.mode,
                                  self.sigma_scale, self.padding_mode, self.cval, self.process_group, self.shard,
                                  self.compute_dtype, self.channels_last, self.return_labels, self._geom_cache,
                                  args, kwargs, host_out=self.device is not None and self.device.type == "cpu",
                                  host_holder=self._host_holder, flip=flip, into=into)
        self.labels = labels
        self.owned_volumes = owned      # indices (into the input batch) of the volumes returned on THIS rank
        return out


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant",
                             sigma_scale=0.125, padding_mode="constant", cval=0.0, sw_device=None, device=None,
                             progress=False, roi_weight_map=None, process_fn=None, buffer_steps=None, buffer_dim=-1,
                             with_coord=False, *args, process_group=None, **kwargs):
    """Functional form with MONAI's signature (``monai/inferers/utils.py:43-64``).  Returns a fresh tensor."""
    if roi_weight_map is not None or process_fn is not None or buffer_steps or with_coord:
        raise NotImplementedError("roi_weight_map / process_fn / buffer_steps / with_coord are not used on this path")
    roi = tuple(roi_size) if isinstance(roi_size, (tuple, list)) else (roi_size,) * 3
    out, _, _ = _run(inputs, roi, int(sw_batch_size), predictor, overlap, getattr(mode, "value", mode), sigma_scale,
                     getattr(padding_mode, "value", padding_mode), cval, process_group, True, None, True, False,
                     _GEOMETRY, args, kwargs, host_out=device is not None and torch.device(device).type == "cpu")
    return out


def _network_dtype(network, fallback: torch.dtype) -> torch.dtype:
    params = getattr(network, "parameters", None)
    if params is not None:
        for p in params():
            if p.is_floating_point():
                return p.dtype
    return fallback


def _run(inputs, roi_size, sw_batch_size, network, overlap, mode, sigma_scale, padding_mode, cval, group, shard,
         compute_dtype, channels_last, return_labels, cache, args, kwargs, host_out: bool = False,
         host_holder: Optional[dict] = None, flip=0, into=None):
    """Returns ``(logits, labels, owned)``: ``logits[i]`` is the stitched fp32 volume ``owned[i]`` (indices into the
    input batch).  Single process: ``owned`` is every volume, i.e. exactly the reference's return value."""
    seq = None
    if isinstance(inputs, (list, tuple)):
        # volumes as a list ([C, D, H, W] or [1, C, D, H, W] each, one shape): with a process group a rank only needs the
        # entries it stitches, the others may be None - nobody has to hold (or pin) the whole cohort
        seq = [None if t is None else (t if t.dim() == 4 else t[0]) for t in inputs]
        first = next((t for t in seq if t is not None), None)
        if first is None or first.dim() != 4 or any(t is not None and t.shape != first.shape for t in seq):
            raise ValueError("a list input needs at least one volume, and all volumes [C, D, H, W] of one shape")
        probe = first[None]
    else:
        probe = inputs
    if probe.dim() != 5:
        raise ValueError("the B200 inferer handles 3D volumes: inputs must be [B, C, D, H, W]")
    nsp = 3
    ov = tuple(overlap) if isinstance(overlap, (tuple, list)) else (overlap,) * nsp
    for o in ov:
        if o < 0 or o >= 1:
            raise ValueError(f"overlap must be >= 0 and < 1, got {overlap}.")
    sg = tuple(sigma_scale) if isinstance(sigma_scale, (tuple, list)) else (sigma_scale,) * nsp
    flip = ops._flip_mask(flip)
    world, rank = 1, 0
    dist = None
    shard_mode = "contiguous" if shard in (True, False, None) else str(shard)
    if shard and (group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = probe.device if probe.is_cuda else _default_device(network)
    batch = len(seq) if seq is not None else inputs.shape[0]
    orig = tuple(probe.shape[2:])
    roi = tuple(int(r) if r and r > 0 else int(o) for r, o in zip(roi_size, orig))
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad = []
    for k in range(nsp - 1, -1, -1):
        diff = max(roi[k] - orig[k], 0)
        pad.extend([diff // 2, diff - diff // 2])
    if flip and any(pad):
        raise NotImplementedError("mirrored passes on volumes smaller than the window (padded input) are not supported")

    key = (size, roi, ov, mode, sg, batch, str(dev), world, rank, sw_batch_size, shard_mode)
    plan = _plan_for(cache, key, lambda: _Plan(size, roi, ov, mode, sg, batch, dev, world, rank, sw_batch_size, shard_mode))
    slot, owned, shared = plan.slot, plan.owned, plan.shared
    gz, gy, gx = plan.fac
    floor = plan.floor
    cur = torch.cuda.current_stream(dev) if dev.type == "cuda" else None

    # bring in only the volumes this rank touches (host input: this is the H2D copy of the end-to-end path)
    in_events = None
    vol = None
    if plan.local_vols:
        if seq is not None:
            parts = seq[plan.lo:plan.hi]
            if any(t is None for t in parts):
                raise ValueError(f"rank {rank} stitches volumes [{plan.lo}, {plan.hi}) and needs all of them in the list")
        else:
            vol = inputs[plan.lo:plan.hi]
            parts = [vol[i] for i in range(vol.shape[0])]
        on_host = not parts[0].is_cuda
        if on_host and parts[0].dtype == torch.float32 and not any(pad) and not flip and plan.tables:
            # streamed H2D: z-slabs in the order the windows need them, one contiguous copy per (volume, channel, slab)
            srcs = [t.contiguous() for t in parts]
            vol = torch.empty((len(srcs),) + tuple(srcs[0].shape), dtype=torch.float32, device=dev)
            cs = _copy_stream(dev)
            cs.wait_stream(cur)
            in_events = {}
            with torch.cuda.stream(cs):
                for v, src in enumerate(srcs):
                    a = 0
                    for si, b in enumerate(plan.slab_ends):
                        for c in range(src.shape[0]):
                            vol[v, c, a:b].copy_(src[c, a:b], non_blocking=True)
                        in_events[(v, si)] = cs.record_event()
                        a = b
            vol.record_stream(cs)
        elif seq is not None:
            vol = torch.stack([t.to(dev, non_blocking=True) for t in parts])
        elif on_host:
            vol = vol.to(dev, non_blocking=True)
        if vol.dtype != torch.float32:
            vol = vol.float()
        if any(pad):
            vol = F.pad(vol, pad, mode=padding_mode, value=cval)
        vol = vol.contiguous()
    dtype = compute_dtype or _network_dtype(network, torch.float32)

    n_owned = len(owned)
    acc = None
    labels = None
    dst = dscale = dadd = None
    if into is not None:
        dst, dscale, dadd = into
        if host_out or world > 1:
            raise NotImplementedError("into= (running mean over mirrored passes) is a single-process, device-resident option")
    # output slabs of volumes that are stitched by this rank alone can leave while later windows run
    stream_out = host_out and not any(pad) and not flip
    streamable = {slot[v] for v in owned if v not in shared} if stream_out else set()
    out_host, out_stream = None, None
    pending: List[Tuple[int, object]] = []        # reduces in flight: (volume, work handle), oldest first
    next_shared = 0                               # index into `shared` of the next reduce this rank issues

    def ensure_acc(k: int):
        nonlocal acc, labels
        if acc is None:
            acc = torch.zeros((len(plan.local_vols), k) + size, dtype=torch.float32, device=dev)
            if return_labels and n_owned:
                labels = torch.empty((n_owned,) + size, dtype=torch.uint8, device=dev)

    def ensure_host():
        nonlocal out_host, out_stream
        if out_host is None:
            out_host = _host_result((n_owned, acc.shape[1]) + size, host_holder)
            out_stream = _copy_stream(dev)
            acc.record_stream(out_stream)

    def finish_shared(v: int, work) -> None:
        """Owner side of a split volume: the reduce has landed -> normalise (+ argmax) -> ship."""
        if work is not None:
            work.wait()                          # the compute stream waits for the collective, the host does not
        if plan.owner[v] != rank:
            return
        i = slot[v]
        ops.sw_finalize(acc[i:i + 1], plan.fin, gz, gy, gx, floor, roi, None if labels is None else labels[i:i + 1],
                        flip=flip)
        if host_out and not any(pad):
            ensure_host()
            out_stream.wait_event(cur.record_event())
            with torch.cuda.stream(out_stream):
                out_host[i].copy_(acc[i], non_blocking=True)

    zeros = None

    def issue_reduces(upto_batch: int, k: int) -> None:
        """Issue, in volume order (identical on every rank), the reduce of every split volume whose share on THIS rank is
        complete after batch ``upto_batch``; each is asynchronous - the next windows run while it is in flight - and the
        owner normalises a volume one reduce later (software pipelining), so nothing ever blocks on the collective."""
        nonlocal next_shared, zeros
        while next_shared < len(shared):
            v = shared[next_shared]
            if plan.done_after[v] > upto_batch:
                break
            if v in slot:
                buf = acc[slot[v]]
            else:                                # a rank that holds no window of v still joins the collective
                if zeros is None:
                    zeros = torch.zeros((k,) + size, dtype=torch.float32, device=dev)
                buf = zeros
            root = plan.owner[v]
            work = dist.reduce(buf, dst=dist.get_global_rank(group, root) if group is not None else root,
                               op=dist.ReduceOp.SUM, group=group, async_op=True)
            pending.append((v, work))
            next_shared += 1
            while len(pending) > 1:
                finish_shared(*pending.pop(0))

    nb = len(plan.tables)
    k_out = getattr(network, "out_chans", None)
    k_out = None if k_out is None else int(k_out)
    # A rank without a single window cannot learn the channel count from its own forward; if the network does not say
    # (`out_chans`), ALL ranks agree on it with one all-reduce after their loops - and, to keep the order of collectives
    # identical everywhere, no reduce is issued before that point.
    agree = world > 1 and bool(shared) and k_out is None and batch * plan.nwin < world
    for j in range(nb):
        if in_events is not None:
            cur.wait_event(in_events[plan.need[j]])   # copies are in order: the last slab suffices
        win = ops.sw_gather(vol, plan.in_tables[j], roi, dtype, channels_last, flip)
        if channels_last:
            win = win.permute(0, 4, 1, 2, 3)  # [n, C, r, r, r] with channels-last-3d strides, zero copy
        seg = network(win, *args, **kwargs)
        if not isinstance(seg, torch.Tensor):
            raise NotImplementedError("the B200 inferer stitches a single tensor output")
        if tuple(seg.shape[2:]) != roi:
            raise NotImplementedError("network output must have the window's spatial size")
        k_out = seg.shape[1]
        ensure_acc(k_out)
        if seg.stride(1) == 1 and seg.shape[1] > 1:
            ops.sw_accumulate(seg.permute(0, 2, 3, 4, 1), acc, plan.tables[j], gz, gy, gx, floor, True, flip)
        else:
            ops.sw_accumulate(seg, acc, plan.tables[j], gz, gy, gx, floor, False, flip)
        for (sl, za, zb) in plan.sched[j]:
            if sl not in streamable:
                continue
            ensure_host()
            ops.sw_finalize(acc[sl:sl + 1], plan.fin, gz, gy, gx, floor, roi,
                            None if labels is None else labels[sl:sl + 1], z_range=(za, zb))
            out_stream.wait_event(cur.record_event())
            with torch.cuda.stream(out_stream):
                for k in range(acc.shape[1]):
                    out_host[sl, k, za:zb].copy_(acc[sl, k, za:zb], non_blocking=True)
        if world > 1 and shared and j < nb - 1 and not agree:
            issue_reduces(j, k_out)

    if world > 1 and shared:
        if agree:
            k_out = _agree_channels(k_out, dist, group, dev)
        if acc is None and (n_owned or any(v in slot for v in shared)):
            ensure_acc(k_out)
        issue_reduces(nb, k_out)
        while pending:
            finish_shared(*pending.pop(0))
    out = None
    if n_owned:
        if acc is None:      # an owner always holds at least one window of its volume
            raise RuntimeError("internal: owned volumes without an accumulator")
        rest = [v for v in owned if slot[v] not in streamable and v not in shared]
        if rest:
            if len(rest) == n_owned and dst is not None:
                ops.sw_finalize(acc[:n_owned], plan.fin, gz, gy, gx, floor, roi, labels, flip=flip, dst=dst,
                                dst_scale=dscale, dst_add=dadd)
            elif len(rest) == n_owned:
                ops.sw_finalize(acc[:n_owned], plan.fin, gz, gy, gx, floor, roi, labels, flip=flip)
            else:
                for v in rest:
                    i = slot[v]
                    ops.sw_finalize(acc[i:i + 1], plan.fin, gz, gy, gx, floor, roi, None if labels is None else labels[i:i + 1],
                                    flip=flip)
        out = acc[:n_owned]
    if out is not None and any(pad):
        crop = [slice(None), slice(None)]
        for sp in range(nsp):
            lo_ = pad[(nsp - 1 - sp) * 2]
            crop.append(slice(lo_, lo_ + orig[sp]))
        out = out[tuple(crop)]
        if labels is not None:
            labels = labels[tuple(crop[1:])]
    if host_out and out is not None:
        if any(pad):
            out_host = _host_result(out.shape, host_holder)
            out_host.copy_(out, non_blocking=True)
        else:
            ensure_host()
            for v in owned:                      # volumes that left neither slab by slab nor after their reduce
                i = slot[v]
                if i not in streamable and v not in shared:
                    out_host[i].copy_(out[i], non_blocking=True)
            out_stream.synchronize()
        cur.synchronize()
        out = out_host
    return out, labels, owned


def _default_device(network) -> torch.device:
    params = getattr(network, "parameters", None)
    if params is not None:
        for p in params():
            return p.device
    return torch.device("cuda", torch.cuda.current_device())


def _agree_channels(k_local: Optional[int], dist, group, dev) -> int:
    """Output channel count for ranks that stitched nothing (they still join the reduce with zeros)."""
    t = torch.tensor([0 if k_local is None else int(k_local)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())
