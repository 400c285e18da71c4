"""Where do the milliseconds between the window forwards go?  Times the stitching kernels of one 4x240x240x155 volume at 6 windows
per forward (gather, accumulate, finalize, accumulator reset) and the whole inferer call against 3 bare graph replays."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waveformer_b200 import ops, prepare_inference  # noqa: E402
from waveformer_b200.graphs import GraphedForward  # noqa: E402
from waveformer_b200.inferers import SlidingWindowInferer  # noqa: E402
from waveformer_b200.network_models import Waveformer  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
m = prepare_inference(Waveformer(img_size=(128,) * 3, patch_size=2, in_chans=4, out_chans=4, depths=[2] * 4, feat_size=[48, 96, 192, 384],
                                 num_heads=[3, 6, 12, 24], drop_path_rate=0.1).eval().to(dev), torch.bfloat16)
g = GraphedForward(m)
vol = torch.randn(1, 4, 240, 240, 155, device=dev)
inf = SlidingWindowInferer(roi_size=(128,) * 3, sw_batch_size=6, overlap=0.5, mode="gaussian", return_labels=True)


def ev(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


with torch.no_grad():
    t_all = ev(lambda: inf(vol, g))
    x = torch.randn(6, 4, 128, 128, 128, device=dev).contiguous(memory_format=torch.channels_last_3d)
    t_fw = ev(lambda: g(x))
    print(f"inferer call: {t_all:.2f} ms per volume; bare graph replay of a 6-window forward: {t_fw:.2f} ms  ->  3 forwards {3 * t_fw:.2f} ms, "
          f"everything else {t_all - 3 * t_fw:.2f} ms")
    t0 = time.perf_counter()
    for _ in range(5):
        inf(vol, g)
    torch.cuda.synchronize()
    print(f"wall clock per call: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
    # the same call under the torch profiler: device time by kernel outside the graph
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        inf(vol, g)
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0 and ("sw_" in e.key or "Memset" in e.key or "memset" in e.key or "fill" in e.key.lower() or "copy" in e.key.lower())]
    for k, t, c in sorted(rows, key=lambda r: -r[1])[:12]:
        print(f"  {t / 1e3:8.3f} ms x{c:3d}  {k[:110]}")
