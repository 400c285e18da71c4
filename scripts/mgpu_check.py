#!/usr/bin/env python
"""Multi-GPU check of the sharded inferer (run under torchrun, one rank per GPU):
  1. ONE 4x240x240x155 volume split over all ranks: window shards + a single NCCL reduce(SUM) of the stitched logit volume
     to rank 0 must equal the unsharded result computed by rank 0 alone (same kernels, so agreement is to fp32 rounding
     of the summation order).
  2. timing of that one-volume case (latency scaling) with CUDA events, max over ranks.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import prepare_inference  # noqa: E402
from waveformer_b200.inferers import SlidingWindowInferer  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4,
                                 feat_size=[48, 96, 192, 384], num_heads=[3, 6, 12, 24]).eval().to(dev), torch.bfloat16)
x = torch.randn((1, 4, 240, 240, 155), generator=torch.Generator().manual_seed(1)).to(dev)
sharded = SlidingWindowInferer(roi_size=(128,) * 3, sw_batch_size=2, overlap=0.5, mode="gaussian", return_labels=True)
alone = SlidingWindowInferer(roi_size=(128,) * 3, sw_batch_size=2, overlap=0.5, mode="gaussian", return_labels=True, shard=False)
with torch.no_grad():
    y = sharded(x, m)
    if rank == 0:
        ref = alone(x, m)
        err = float((y - ref).abs().max() / ref.abs().max())
        agree = float((sharded.labels == alone.labels).float().mean())
        print(f"one volume over {world} ranks vs rank 0 alone: max-rel {err:.2e}, label agreement {agree:.6f}", flush=True)
        assert err < 1e-5 and agree > 0.99999
    else:
        assert y is None
    for it in range(2):
        for _ in range(2):
            sharded(x, m)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sharded(x, m)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"one volume, {world} GPUs: {float(t):.1f} ms per volume ({240 * 240 * 155 / float(t) / 1e3:.1f} M voxels/s)", flush=True)
dist.barrier()
dist.destroy_process_group()
