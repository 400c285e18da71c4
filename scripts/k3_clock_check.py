"""Cross-check of the rolling-row convolution's device time: CUDA-event time of graph-replayed launches (production and
diagnostic instantiation) against the in-kernel clock64 totals, with the SM clock sampled by nvidia-smi while the kernel loops."""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waveformer_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(2, 128, 128, 128, 48, generator=g).half().to(dev).permute(0, 4, 1, 2, 3)
w = (torch.randn(48, 48, 3, 3, 3, generator=g) / 36).half().to(dev)
clk = torch.zeros(148, 3, 8, dtype=torch.int64, device=dev)


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(iters):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gr.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


samples = []
stop = False


def sampler():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                             capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.05)


with torch.no_grad():
    for iters in (1, 10):
        t = timed(lambda: ops.conv3d_k3_c48(x, w), iters)
        print(f"production kernel, {iters:3d} launches per graph: {t:.1f} us per launch")
    t = timed(lambda: ops.conv3d_k3_c48(x, w, stage_clocks=clk), 10)
    c = clk.double().cpu()[:, :, 7]
    print(f"diagnostic kernel, 10 launches per graph: {t:.1f} us per launch; in-kernel clocks per CTA: mean {c.mean():.0f}, max {c.max():.0f}, min {c.min():.0f}"
          f"  -> {c.max() / t / 1e3:.2f} GHz implied")
    y = torch.nn.functional.conv3d
    t = timed(lambda: y(x, w, padding=1), 10)
    print(f"cudnn 48->48: {t:.1f} us")
    # who are the stragglers?  wall clock (globaltimer) against SM clocks (clock64) per CTA
    for rep in range(3):
        clk.zero_()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(1 if rep == 0 else 6):
            ops.conv3d_k3_c48(x, w, stage_clocks=clk)
        b_.record()
        torch.cuda.synchronize()
        c = clk.cpu()
        tot = c[:, 1, 7].double()
        t_in, t_out, smid = c[:, 0, 6], c[:, 1, 6], c[:, 2, 6]
        wall = (t_out - t_in).double()
        print(f"run {rep} ({1 if rep == 0 else 6} launches, {a_.elapsed_time(b_) * 1e3:.0f} us by events; last launch): grid wall clock {int(t_out.max() - t_in.min()) / 1e3:.1f} us, "
              f"entry spread {int(t_in.max() - t_in.min()) / 1e3:.1f} us, per-CTA wall min / mean / max {wall.min() / 1e3:.1f} / {wall.mean() / 1e3:.1f} / {wall.max() / 1e3:.1f} us, "
              f"clocks min / mean / max {tot.min():.0f} / {tot.mean():.0f} / {tot.max():.0f}, clocks per ns {float((tot / wall).min()):.2f} .. {float((tot / wall).max()):.2f}")
        order = torch.argsort(wall, descending=True)
        for i in order[:6].tolist():
            print(f"  cta {i:3d} on sm {int(smid[i]):3d}: wall {wall[i] / 1e3:7.1f} us, clocks {int(tot[i]):8d}; issuer wait acc / wait row / mma: {int(c[i,1,0]):7d} {int(c[i,1,1]):7d} {int(c[i,1,2]):7d}")
