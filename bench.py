#!/usr/bin/env python
"""Benchmark of the WaveFormer hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference algorithm on the host CPU cores

Metric (BASELINE.json): BraTS 4-channel sliding-window voxels/s - one step = sliding-window inference (ROI 128^3,
overlap 0.5, gaussian blending, sw_batch 2, bf16) over V synthetic 4x240x240x155 volumes, V = N GPUs (weak scaling:
18 windows per GPU per step, windows sharded over one process per GPU).  `value` is measured with the volumes
resident in HBM; `e2e` goes through the public inferer call with PINNED HOST volumes (H2D inside the timed region) and
reads the stitched fp32 logits back to the host.  `roofline` is the Haar DWT kernel (north star's second metric,
"DWT/IDWT HBM GB/s") on BASELINE config 2 (2x48x128^3), timed live with CUDA events.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL = (4, 240, 240, 155)
VOXELS = 240 * 240 * 155
ROI = (128, 128, 128)
WINDOWS_PER_VOLUME = 18
MODEL_KW = dict(img_size=ROI, patch_size=2, in_chans=4, out_chans=4, depths=[2, 2, 2, 2], feat_size=[48, 96, 192, 384],
                num_heads=[3, 6, 12, 24], drop_path_rate=0.1)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks ---------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(power), samples=len(sm),
                    reasons=sorted(reasons))


# ------------------------------------------------------------------------------------------------ helpers --------
def event_ms(fn, iters: int, warm: int = 3) -> float:
    """Average device time of fn() over `iters` launches, CUDA events on the current stream, sync on both sides."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def kernel_rooflines(pk):
    """Per-kernel achieved bandwidth on BASELINE config 2 (2x48x128^3), bf16 and fp32; inputs (403 / 805 MB) are larger
    than the 126 MB L2, so consecutive launches cannot hit in cache."""
    from waveformer_b200 import ops
    out = {}
    for name, dtype, esz in (("bf16", torch.bfloat16, 2), ("f32", torch.float32, 4)):
        x = torch.randn((2, 48, 128, 128, 128), device="cuda", dtype=torch.float32).to(dtype)
        n = x.numel()
        alg = 2 * n * esz  # read N, write N (LL N/8 + 7 details N/8)
        ll, hf = ops._dwt_ncdhw_raw(x, True)
        t_dwt = event_ms(lambda: ops._dwt_ncdhw_raw(x, True), 20)
        t_idwt = event_ms(lambda: ops._idwt_ncdhw_raw(ll, hf), 20)
        xc = x.view(2, 128, 128, 128, 48)  # same bytes read as a channels-last volume
        llc, hfc = ops._dwt_ndhwc_raw(xc, True)
        t_dwtc = event_ms(lambda: ops._dwt_ndhwc_raw(xc, True), 20)
        t_idwtc = event_ms(lambda: ops._idwt_ndhwc_raw(llc, hfc), 20)
        for k, t in (("dwt3d_ncdhw", t_dwt), ("idwt3d_ncdhw", t_idwt), ("dwt3d_ndhwc", t_dwtc), ("idwt3d_ndhwc", t_idwtc)):
            gbs = alg / (t * 1e-3) / 1e9
            out[f"{k}_{name}"] = dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm_gbs"], unit="GB/s",
                                      frac=round(gbs / pk["hbm_gbs"], 4), ms=round(t, 4), algorithmic_bytes=alg)
        out[f"roundtrip_ncdhw_{name}"] = dict(bound="hbm", achieved=round(2 * alg / ((t_dwt + t_idwt) * 1e-3) / 1e9, 1),
                                              peak=pk["hbm_gbs"], unit="GB/s",
                                              frac=round(2 * alg / ((t_dwt + t_idwt) * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                                              ms=round(t_dwt + t_idwt, 4), algorithmic_bytes=2 * alg)
        del x, ll, hf, llc, hfc, xc
        torch.cuda.empty_cache()
    # window attention: stage-1 level-1 geometry at sw_batch 2 (128 windows of 512 tokens, C=48, 3 heads), bf16
    from waveformer_b200.network_models import Attention
    # (the inference policy's configuration: fp32 stream in, fp16 tcgen05 operands, fp32 out; 3 launches per call)
    att = Attention(48, num_heads=3, qkv_bias=True, window_size=8).cuda().eval()
    att.compute_dtype, att.out_dtype = torch.float16, torch.float32
    xa = torch.randn((2, 32, 32, 32, 48), device="cuda")
    with torch.no_grad():
        t = event_ms(lambda: att.forward_grid(xa), 20)
    flops = 128 * (4096 * 48 ** 2 + 1048576 * 48)
    tf = flops / (t * 1e-3) / 1e12
    out["window_attention_c48_tc"] = dict(bound="tensor", achieved=round(tf, 2), peak=pk["bf16_tflops"], unit="TFLOP/s",
                                          frac=round(tf / pk["bf16_tflops"], 5), ms=round(t, 4), algorithmic_flops=flops,
                                          note="head_dim 16: exponent-bound (65536 ex2 per 128x512 tile vs 512 tensor clk), DESIGN.md 5")
    # 3x3x3 convolution 48 -> 48 on 2 x 128^3 (conv2 of encoder1 / decoder1): the largest single kernel of the window forward
    xk = torch.randn((2, 128, 128, 128, 48), device="cuda").bfloat16().permute(0, 4, 1, 2, 3)
    wk = (torch.randn((48, 48, 3, 3, 3), device="cuda") / 36).bfloat16()
    with torch.no_grad():
        t = event_ms(lambda: ops.conv3d_k3_c48(xk, wk), 10)
    flops = 2 * xk.numel() * 48 * 27
    tf = flops / (t * 1e-3) / 1e12
    out["conv3d_k3_c48_tc"] = dict(bound="tensor", achieved=round(tf, 2), peak=pk["bf16_tflops"], unit="TFLOP/s",
                                   frac=round(tf / pk["bf16_tflops"], 5), ms=round(t, 4), algorithmic_flops=flops,
                                   note="N = 48 output channels: one tcgen05.mma per 128 x 16 activation operand, DESIGN.md 4")
    del xk, wk
    torch.cuda.empty_cache()
    return out


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_patch_seconds(steps: int, warmup: int):
    """The reference algorithm (oracle port, fp32) on the host cores: one 128^3 window forward per step."""
    host_threads()
    from oracle.model import waveformer_forward
    from oracle.state import ModelConfig, make_state_dict
    cfg = ModelConfig(img_size=ROI)
    sd = make_state_dict(cfg, seed=0)
    x = torch.randn((1,) + (4,) + ROI, generator=torch.Generator().manual_seed(0))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            waveformer_forward(sd, x, cfg)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: /root/reference is Python and
    does not travel to the GPU box), all host threads, bounded sample = one 128^3 window forward per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    sec = cpu_patch_seconds(args.steps, args.warmup)
    value = VOXELS / (WINDOWS_PER_VOLUME * sec)
    sample = "1 of 18 windows per step: one 1x4x128^3 fp32 forward of the oracle port; volume time = 18 x patch time"
    line = dict(impl="reference", metric="sliding_window_voxels_per_s", value=value, unit="voxels/s", n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=sec * 1e3 * WINDOWS_PER_VOLUME, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=workload_config(1),
                cpu_baseline=dict(value=value, unit="voxels/s", cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit="voxels/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(volumes: int, sw_batch: int = 2):
    return dict(workload="WaveFormer sliding-window inference, synthetic 4x240x240x155 volume(s), ROI 128^3, overlap 0.5, "
                         f"gaussian blending, sw_batch_size {sw_batch} (BASELINE configs[2]; configs[3] at N>1)",
                volumes_per_step=volumes, windows_per_volume=WINDOWS_PER_VOLUME, roi=list(ROI), overlap=0.5,
                blend="gaussian", sw_batch_size=sw_batch, parallelism="windows sharded over one process per GPU",
                l2_policy="inputs and activations (>= 143 MB per volume) exceed the 126 MB L2; no explicit flush",
                launch="window forward replayed as a CUDA graph (waveformer_b200.graphs.GraphedForward)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--volumes", type=int, default=0, help="volumes per step (default: one per GPU)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch every kernel of the window forward eagerly")
    ap.add_argument("--sw-batch", type=int, default=2, help="windows per forward (the reference's 4_predict.py uses 2)")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch.distributed as dist
    from waveformer_b200 import ops
    from waveformer_b200.inferers import SlidingWindowInferer
    from waveformer_b200.network_models import Waveformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    volumes = args.volumes or world
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    pk = peaks()

    torch.backends.cudnn.benchmark = True   # let cuDNN time its algorithms for the (fixed) window shapes during warm-up
    torch.manual_seed(0)  # identical random-init weights on every rank
    from waveformer_b200 import prepare_inference
    model = prepare_inference(Waveformer(**MODEL_KW).eval().to(dev), dtype)   # bf16 = the documented precision policy
    if not args.no_cuda_graph:
        from waveformer_b200.graphs import GraphedForward
        model = GraphedForward(model)      # the window forward (~450 launches) is replayed as one CUDA graph
    host = torch.randn((volumes,) + VOL, generator=torch.Generator().manual_seed(1)).pin_memory()
    resident = host.to(dev)
    inferer = SlidingWindowInferer(roi_size=ROI, sw_batch_size=args.sw_batch, overlap=0.5, mode="gaussian", return_labels=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        with torch.no_grad():
            return inferer(resident, model)

    # host in, host out: `device="cpu"` is MONAI's argument for where the stitched output lives; the inferer streams the
    # pinned input in and the normalised logits out in z-slabs (inferers.py), and returns when the host copy is complete
    inferer_e2e = SlidingWindowInferer(roi_size=ROI, sw_batch_size=args.sw_batch, overlap=0.5, mode="gaussian",
                                       return_labels=True, device="cpu")

    def step_e2e():
        with torch.no_grad():
            y = inferer_e2e(host, model)       # H2D of this rank's volumes and D2H of its logits happen inside the call
        assert y is None or not y.is_cuda
        return y

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = volumes * VOXELS / (ms_step * 1e-3)

    # ---- end to end through the public call, host buffers -----------------------------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = None
    for _ in range(args.steps):
        y = step_e2e()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    owned = len(inferer_e2e.owned_volumes)
    touched = max(1, -(-volumes // world)) if volumes >= world else 1
    h2d = touched * 4 * VOXELS * 4                     # fp32 volumes this rank copies in (rank 0's share)
    d2h = owned * 4 * VOXELS * 4                       # fp32 stitched logits this rank reads back
    e2e = dict(value=volumes * VOXELS / (ms_e2e * 1e-3), unit="voxels/s", ms_per_step=ms_e2e,
               h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel rooflines (rank 0, N = 1 view of the kernels) ----------------------------------------------
    kernels = {} if args.no_kernel_rooflines else kernel_rooflines(pk)
    # headline roofline: the DWT3D -> IDWT3D round trip of BASELINE configs[1] (the metric's "DWT/IDWT HBM GB/s"), two
    # launches; `traffic` = DRAM bytes of the same two launches from the committed ncu capture (profiles/)
    roof = dict(kernels.get("roundtrip_ncdhw_bf16", dict(bound="hbm", achieved=None, peak=pk["hbm_gbs"], unit="GB/s", frac=None)))
    roof["kernel"] = "dwt_ncdhw_vec_kernel<bf16> + idwt_ncdhw_vec_kernel<bf16> on 2x48x128^3 (BASELINE configs[1]), per round trip"
    roof["peak_source"] = pk["source"]
    roof["traffic"] = None
    tpath = os.path.join(ROOT, "profiles", "r01_kernel_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tr = json.load(f)
        roof["traffic"] = tr["dwt3d_ncdhw_bf16"]["dram_traffic_bytes"] + tr["idwt3d_ncdhw_bf16"]["dram_traffic_bytes"]
        roof["traffic_source"] = "profiles/r01_kernel_traffic.json (ncu --set full; the last ~50 MB of each launch's writes are still in L2)"
    # the kernel that takes the largest share of the window forward (profiles/r01_launch_summary.txt): the 3^3 tensor-core
    # convolution, two launches per forward (one with the fused input normalisation), tensor-pipe bound
    dominant = None
    if "conv3d_k3_c48_tc" in kernels:
        dominant = dict(kernels["conv3d_k3_c48_tc"])
        dominant["kernel"] = "conv3d_k3_c48_kernel on 2x48x128^3 (conv2 of encoder1 / decoder1)"
        dominant["share_of_forward"] = "2 launches, ~15 % of the window forward's device time (ncu launch list)"
        dominant["traffic"] = 770605000      # dram bytes per launch, profiles/r01_ncu_k3.json (algorithmic 805.3 MB)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sec = cpu_patch_seconds(1, 1)
        cpu = dict(value=VOXELS / (WINDOWS_PER_VOLUME * sec), unit="voxels/s", cores=host_threads(), kind="port",
                   sample="1 of 18 windows: one 1x4x128^3 fp32 forward of the oracle port (1 warm-up + 1 timed), "
                          "volume time = 18 x patch time", patch_seconds=sec)
    line = dict(metric="sliding_window_voxels_per_s", value=value, unit="voxels/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.dtype, data="synthetic", config=workload_config(volumes, args.sw_batch), clocks=clocks, e2e=e2e,
                gpu_launches=launches, roofline=roof, roofline_dominant_by_time=dominant, roofline_kernels=kernels,
                cpu_baseline=cpu)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
