#!/usr/bin/env python
"""Benchmark of the WaveFormer hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference algorithm on the host CPU cores

Metric (BASELINE.json): BraTS 4-channel sliding-window voxels/s - one step = sliding-window inference (ROI 128^3,
overlap 0.5, gaussian blending, 16-bit precision policy; 6 windows per forward, the reference script's 2 beside it as
`sw_batch_2`) over V synthetic 4x240x240x155 volumes:
  N = 1  V = 1  (BASELINE configs[2]);
  N > 1  V = 64 (BASELINE configs[3]): the 1152 windows are sharded over one process per GPU in contiguous runs cut at
         window granularity; a volume whose windows land on several ranks costs one NCCL reduce of its stitched logits.
         The same line carries `strong`: ONE volume split over all N ranks (2-3 windows each, one reduce; latency) and 8
         volumes dealt window by window over all ranks (every volume reduced while the next one's windows run).
`value` is measured with the volumes resident in HBM; `e2e` goes through the public inferer call with PINNED HOST volumes
(H2D inside the timed region) and reads the stitched fp32 logits back to the host.  `roofline` is the Haar DWT kernel
(north star's second metric, "DWT/IDWT HBM GB/s") on BASELINE config 2 (2x48x128^3), timed live with CUDA events.
At N = 1 the line also carries `parity` (16-bit policy vs the fp32 path on this run's weights), `train_step` (BASELINE
configs[4]: bf16-autocast forward + backward + AdamW at 2x4x128^3) and `tta8` (8-pass mirror TTA folded into the stitching
kernels).
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
    # (the inference policy's configuration: fp32 stream in, error-compensated fp16 tcgen05 operands - 3 MMAs per product
    # in the projections and the scores - fp32 out; 3 launches per call; FLOPs counted once, as the algorithm needs them)
    att = Attention(48, num_heads=3, qkv_bias=True, window_size=8).cuda().eval()
    att.compute_dtype, att.out_dtype, att.split_operands = torch.float16, torch.float32, True
    xa = torch.randn((2, 32, 32, 32, 48), device="cuda")
    with torch.no_grad():
        t = event_ms(lambda: att.forward_grid(xa), 20)
    flops = 128 * (4096 * 48 ** 2 + 1048576 * 48)
    tf = flops / (t * 1e-3) / 1e12
    out["window_attention_c48_tc"] = dict(bound="tensor", achieved=round(tf, 2), peak=pk["bf16_tflops"], unit="TFLOP/s",
                                          frac=round(tf / pk["bf16_tflops"], 5), ms=round(t, 4), algorithmic_flops=flops,
                                          note="head_dim 16: exponent-bound (65536 ex2 per 128x512 tile vs 512 tensor clk), DESIGN.md 5")
    # 3x3x3 convolution 48 -> 48 on 2 x 128^3 (conv2 of encoder1 / decoder1; twice more as the two passes of decoder1.conv1):
    # the largest single kernel of the window forward.  fp16 = the 16-bit policy's storage format.
    xk = torch.randn((2, 128, 128, 128, 48), device="cuda").half().permute(0, 4, 1, 2, 3)
    wk = (torch.randn((48, 48, 3, 3, 3), device="cuda") / 36).half()
    with torch.no_grad():
        t = event_ms(lambda: ops.conv3d_k3_c48(xk, wk), 10)
    flops = 2 * xk.numel() * 48 * 27
    tf = flops / (t * 1e-3) / 1e12
    out["conv3d_k3_c48_tc"] = dict(bound="tensor", achieved=round(tf, 2), peak=pk["bf16_tflops"], unit="TFLOP/s",
                                   frac=round(tf / pk["bf16_tflops"], 5), ms=round(t, 4), algorithmic_flops=flops,
                                   note="rolling-row kernel: 27 tcgen05.mma 128x144x16 per output row (floor 72 clk each), TMA row loads, "
                                        "statistics + bulk store in the epilogue, DESIGN.md 4")
    x96 = torch.randn((2, 128, 128, 128, 96), device="cuda").half().permute(0, 4, 1, 2, 3)
    w96 = (torch.randn((48, 96, 3, 3, 3), device="cuda") / 50).half()
    with torch.no_grad():
        t = event_ms(lambda: ops.conv3d_k3_c96_c48(x96, w96), 10)
    flops = 2 * xk.numel() * 96 * 27
    tf = flops / (t * 1e-3) / 1e12
    out["conv3d_k3_c96_c48_tc"] = dict(bound="tensor", achieved=round(tf, 2), peak=pk["bf16_tflops"], unit="TFLOP/s",
                                       frac=round(tf / pk["bf16_tflops"], 5), ms=round(t, 4), algorithmic_flops=flops,
                                       note="decoder1.conv1 on the concatenation buffer: two passes of the rolling-row kernel, the second "
                                            "adds the first's 16-bit result to its accumulators")
    del x96, w96
    # InstanceNorm + LeakyReLU pass of the 128^3 residual blocks on the bench's six windows (1.2 GB in + 1.2 GB out, fp16): the
    # register-constant, software-pipelined kernel; and the output head (InstanceNorm + InstanceNorm'd residual + LeakyReLU + 1^3 conv)
    xi = torch.randn((6, 128, 128, 128, 48), device="cuda").half().permute(0, 4, 1, 2, 3)
    sti = ops.instance_norm_stats(xi)
    with torch.no_grad():
        t = event_ms(lambda: ops.instance_norm_act(xi, "leakyrelu", 0.01, stats=sti), 10)
    alg = 2 * xi.numel() * 2
    gbs = alg / (t * 1e-3) / 1e9
    out["instnorm_apply_f16"] = dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm_gbs"], unit="GB/s", frac=round(gbs / pk["hbm_gbs"], 4),
                                     ms=round(t, 4), algorithmic_bytes=alg, note="6 x 48 x 128^3: one read + one write, DESIGN.md 4")
    wh, bh = torch.randn((4, 48, 1, 1, 1), device="cuda") / 7, torch.randn(4, device="cuda") * 0.1
    with torch.no_grad():
        ri = torch.randn((6, 128, 128, 128, 48), device="cuda").half().permute(0, 4, 1, 2, 3)
        t = event_ms(lambda: ops.instance_norm_act_head(xi, wh, bh, "leakyrelu", 0.01, res=ri, res_norm=True, stats=sti, res_stats=sti), 10)
    alg = 2 * xi.numel() * 2 + 6 * 128 ** 3 * 4 * 4
    gbs = alg / (t * 1e-3) / 1e9
    out["instnorm_apply_head_f16"] = dict(bound="hbm", achieved=round(gbs, 1), peak=pk["hbm_gbs"], unit="GB/s",
                                          frac=round(gbs / pk["hbm_gbs"], 4), ms=round(t, 4), algorithmic_bytes=alg,
                                          note="6 x 48 x 128^3 activation + residual in, 4 fp32 logits per voxel out")
    del xi, sti, ri
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
    times.sort()
    return times[len(times) // 2], times


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: /root/reference is Python and
    does not travel to the GPU box), all host threads, bounded sample = one 128^3 window forward per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    sec, times = cpu_patch_seconds(max(args.steps, 3), max(args.warmup, 1))
    value = VOXELS / (WINDOWS_PER_VOLUME * sec)
    sample = (f"1 of 18 windows per step: one 1x4x128^3 fp32 forward of the oracle port, median of {len(times)} timed steps "
              f"on {cores} host threads; volume time = 18 x patch time")
    line = dict(impl="reference", metric="sliding_window_voxels_per_s", value=value, unit="voxels/s", n_gpus=args.gpus,
                steps=len(times), warmup=max(args.warmup, 1), ms_per_step=sec * 1e3 * WINDOWS_PER_VOLUME, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=workload_config(1, 1),
                cpu_baseline=dict(value=value, unit="voxels/s", cores=cores, kind="port", sample=sample,
                                  patch_seconds=[round(t, 4) for t in times]),
                e2e=dict(value=value, unit="voxels/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def workload_config(volumes: int, world: int, sw_batch: int = 6):
    which = "BASELINE configs[2]" if world == 1 else "BASELINE configs[3]"
    return dict(workload=f"WaveFormer sliding-window inference, {volumes} synthetic 4x240x240x155 volume(s) per step, ROI 128^3, "
                         f"overlap 0.5, gaussian blending, sw_batch_size {sw_batch} ({which})",
                volumes_per_step=volumes, windows_per_volume=WINDOWS_PER_VOLUME, roi=list(ROI), overlap=0.5,
                blend="gaussian", sw_batch_size=sw_batch,
                parallelism=("one GPU" if world == 1 else
                             f"{volumes * WINDOWS_PER_VOLUME} windows sharded over {world} processes (one per GPU) in contiguous "
                             "window-granular runs; NCCL reduce of the stitched logits for every volume split across ranks"),
                precision="16-bit policy (waveformer_b200.prepare_inference): fp16 storage / tensor-core operands, fp32 "
                          "accumulation, fp32 encoder stream, error-compensated fp16 attention operands, fp32 logits",
                l2_policy="inputs and activations (>= 143 MB per volume) exceed the 126 MB L2; no explicit flush",
                launch="window forward replayed as a CUDA graph (waveformer_b200.graphs.GraphedForward)")


def timed(fn, steps, barrier, max_over_ranks):
    """K calls of fn() between two CUDA events, barrier + synchronize on both sides; returns ms per call (max over ranks)."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for _ in range(steps):
        out = fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps, out


def parity_probe(dev, model16):
    """16-bit policy vs the fp32 product path (itself <= 1e-5 from the CPU oracle, tests/test_gpu_model.py) on one
    1x4x128^3 window with THIS run's weights (the constructors' random initialisation, seed 0)."""
    from waveformer_b200 import prepare_inference
    from waveformer_b200.network_models import Waveformer
    torch.manual_seed(0)
    m32 = prepare_inference(Waveformer(**MODEL_KW).eval().to(dev), torch.float32)
    x = torch.randn((1, 4) + ROI, generator=torch.Generator().manual_seed(2)).to(dev)
    with torch.no_grad():
        ref = m32(x).float()
        y = model16(x).float()
    err = float((y - ref).abs().max() / ref.abs().max())
    agree = float((y.argmax(1) == ref.argmax(1)).float().mean())
    del m32
    torch.cuda.empty_cache()
    return dict(reference="fp32 product path on the same weights and window", max_rel_logit_error=err, argmax_agreement=agree,
                gates=dict(max_rel_logit_error=2e-2, argmax_agreement=0.999),
                note="unit-gain stress weights of the test suite: 99.90 % (tests/test_gpu_model.py, DESIGN.md 6)")


def train_step_probe(dev, steps=3, warmup=2):
    """BASELINE configs[4]: one training step at 2x4x128^3 - bf16 autocast forward, Dice + CE loss, backward through the
    custom DWT / attention / IDWT kernels, fused AdamW - timed with CUDA events."""
    from waveformer_b200 import ops
    from waveformer_b200.losses import DiceCELoss
    from waveformer_b200.network_models import Waveformer
    torch.manual_seed(0)
    m = Waveformer(**MODEL_KW).to(dev).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True)
    loss_fn = DiceCELoss(to_onehot_y=True, softmax=True)
    x = torch.randn((2, 4) + ROI, device=dev)
    y = torch.randint(0, 4, (2, 1) + ROI, device=dev)
    torch.cuda.reset_peak_memory_stats(dev)
    fb = []

    def step():
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = m(x)
        loss = loss_fn(logits, y)
        loss.backward()
        b.record()
        opt.step()
        c.record()
        return loss, a, b, c

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    l0 = ops.LAUNCHES
    rec = [step() for _ in range(steps)]
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(c) for _, a, _, c in rec) / steps
    ms_fb = sum(a.elapsed_time(b) for _, a, b, _ in rec) / steps
    out = dict(config="BASELINE configs[4]: batch 2x4x128^3, DiceCELoss(to_onehot_y, softmax), bf16 autocast, fused AdamW",
               ms_per_step=ms, ms_forward_backward=ms_fb, samples_per_s=2 / (ms * 1e-3), dtype="bf16 autocast (fp32 master weights)",
               loss=float(rec[-1][0].detach()), peak_gb=torch.cuda.max_memory_allocated(dev) / 2 ** 30,
               own_kernel_launches_per_step=(ops.LAUNCHES - l0) // steps, steps=steps, warmup=warmup)
    del m, opt, rec
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--volumes", type=int, default=0, help="volumes per step (default: 1 on one GPU, 64 on several)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"], help="bf16 = the 16-bit precision policy")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch every kernel of the window forward eagerly")
    ap.add_argument("--sw-batch", type=int, default=6,
                    help="windows per forward.  A tuning knob of the inferer, not part of the workload: windows are independent and every "
                         "normalisation is per sample, so the stitched result does not depend on it.  6 = three forwards per volume; the "
                         "reference's 4_predict.py passes 2, reported beside the headline as `sw_batch_2`")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip parity / train_step / tta8 / strong")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch.distributed as dist
    from waveformer_b200 import ops
    from waveformer_b200.inferers import SlidingWindowInferer, shard_windows
    from waveformer_b200.network_models import Waveformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    volumes = args.volumes or (1 if world == 1 else 64)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    pk = peaks()

    torch.backends.cudnn.benchmark = True   # let cuDNN time its algorithms for the (fixed) window shapes during warm-up
    torch.manual_seed(0)  # identical random-init weights on every rank
    from waveformer_b200 import prepare_inference
    eager = prepare_inference(Waveformer(**MODEL_KW).eval().to(dev), dtype)   # bf16 = the documented 16-bit policy
    model = eager
    if not args.no_cuda_graph:
        from waveformer_b200.graphs import GraphedForward
        model = GraphedForward(eager)      # the window forward (~400 launches) is replayed as one CUDA graph
    # the same synthetic cohort on every rank (device generator, fixed seed); a rank only reads the volumes it stitches
    resident = torch.randn((volumes,) + VOL, generator=torch.Generator(device=dev).manual_seed(1), device=dev)
    kw = dict(roi_size=ROI, sw_batch_size=args.sw_batch, overlap=0.5, mode="gaussian", return_labels=True)
    inferer = SlidingWindowInferer(**kw)

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

    # ---- device-resident throughput -------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCHES
    ms_step, _ = timed(step_resident, args.steps, barrier, max_over_ranks)
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = volumes * VOXELS / (ms_step * 1e-3)

    # ---- end to end through the public call, host buffers -----------------------------------------------------
    # host in, host out: `device="cpu"` is MONAI's argument for where the stitched output lives; the inferer streams the
    # pinned input in and the normalised logits out in z-slabs (inferers.py) and returns when the host copy is complete.
    # Volumes are handed over as a list in which a rank holds (pinned) only the volumes it stitches; reuse_output=True
    # recycles the pinned result buffer (pinning 143 MB per volume costs more than moving it).
    total_w = volumes * WINDOWS_PER_VOLUME
    mine = sorted({i // WINDOWS_PER_VOLUME for i in shard_windows(total_w, rank, world)})
    pinned = torch.empty((len(mine),) + VOL, dtype=torch.float32, pin_memory=True)
    for j, v in enumerate(mine):
        pinned[j].copy_(resident[v])
    torch.cuda.synchronize()
    host_list = [None] * volumes
    for j, v in enumerate(mine):
        host_list[v] = pinned[j]
    inferer_e2e = SlidingWindowInferer(device="cpu", reuse_output=True, **kw)

    def step_e2e():
        with torch.no_grad():
            y = inferer_e2e(host_list, model)       # H2D of this rank's volumes and D2H of its logits happen inside the call
        assert y is None or not y.is_cuda
        return y

    e2e_steps = args.steps if world == 1 else min(args.steps, 5)
    for _ in range(2 if world == 1 else 1):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, e2e_steps, barrier, max_over_ranks)
    owned = len(inferer_e2e.owned_volumes)
    h2d = len(mine) * 4 * VOXELS * 4                   # fp32 volumes this rank copies in (rank 0's share)
    d2h = owned * 4 * VOXELS * 4                       # fp32 stitched logits this rank reads back
    e2e = dict(value=volumes * VOXELS / (ms_e2e * 1e-3), unit="voxels/s", ms_per_step=ms_e2e, steps=e2e_steps,
               h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
               api="SlidingWindowInferer(device='cpu', reuse_output=True)(list of pinned host volumes, model)")
    del pinned, host_list, inferer_e2e

    # ---- the reference prediction script's batching (4_predict.py:199-205: sw_batch_size = 2), for continuity with round 1 ----
    sw2 = None
    if world == 1 and args.sw_batch != 2 and not args.no_extras:
        inf2 = SlidingWindowInferer(**dict(kw, sw_batch_size=2))

        def step2():
            with torch.no_grad():
                return inf2(resident, model)

        for _ in range(3):
            step2()
        ms2, _ = timed(step2, args.steps, barrier, max_over_ranks)
        sw2 = dict(value=volumes * VOXELS / (ms2 * 1e-3), unit="voxels/s", ms_per_step=ms2, steps=args.steps, sw_batch_size=2)
        del inf2

    # ---- N > 1: one volume split over all ranks (latency) and every volume split over all ranks (interleaved) ----
    strong = None
    if world > 1 and not args.no_extras:
        one = resident[:1]
        inf1 = SlidingWindowInferer(**kw)
        per_rank = [len(shard_windows(WINDOWS_PER_VOLUME, r, world)) for r in range(world)]

        def step_one():
            with torch.no_grad():
                return inf1(one, model)

        for _ in range(3):
            step_one()
        ms_one, _ = timed(step_one, 10, barrier, max_over_ranks)
        nv = min(8, volumes)
        some = resident[:nv]
        inf8 = SlidingWindowInferer(shard="interleaved", **kw)

        def step_inter():
            with torch.no_grad():
                return inf8(some, model)

        for _ in range(2):
            step_inter()
        ms_inter, _ = timed(step_inter, 5, barrier, max_over_ranks)
        buf = torch.zeros((4,) + VOL[1:], dtype=torch.float32, device=dev)

        def one_reduce():
            dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)

        for _ in range(3):
            one_reduce()
        ms_red, _ = timed(one_reduce, 10, barrier, max_over_ranks)
        # forward time by batch size on this rank (CUDA graph replays), to name the limiter from measurement
        fw = {}
        for b in sorted({min(args.sw_batch, max(per_rank)), max(1, max(per_rank) % args.sw_batch or args.sw_batch)}):
            xb = torch.randn((b, 4) + ROI, device=dev).contiguous(memory_format=torch.channels_last_3d)
            with torch.no_grad():
                fw[b] = event_ms(lambda: model(xb), 5)
        chunks = [min(args.sw_batch, max(per_rank) - s) for s in range(0, max(per_rank), args.sw_batch)]
        compute = sum(fw.get(c, fw[max(fw)] * c / max(fw)) for c in chunks)
        strong = dict(
            one_volume=dict(ms_per_volume=ms_one, voxels_per_s=VOXELS / (ms_one * 1e-3), windows_per_rank=per_rank,
                            max_windows_per_rank=max(per_rank), sharding="contiguous, window granularity", reduces_per_volume=1),
            interleaved=dict(volumes=nv, ms_per_volume=ms_inter / nv, voxels_per_s=nv * VOXELS / (ms_inter * 1e-3),
                             sharding="window g -> rank g % N: every volume split over all ranks, its reduce issued "
                                      "asynchronously while the next volume's windows run", reduces_per_step=nv),
            reduce_ms=ms_red, reduce_bytes=buf.numel() * 4, reduce_gbs=buf.numel() * 4 / (ms_red * 1e-3) / 1e9,
            forward_ms_by_batch={str(k): v for k, v in fw.items()},
            limiter=(f"window forwards on the rank with the most windows: {max(per_rank)} windows = {compute:.1f} ms of the "
                     f"{ms_one:.1f} ms one-volume latency (reduce of the 143 MB logit volume alone: {ms_red:.2f} ms)"))
        del buf, inf1, inf8

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
    # the kernel that takes the largest share of the window forward (profiles/): the 3^3 tensor-core convolution, two
    # launches per forward (one with the fused input normalisation), tensor-pipe bound
    dominant = None
    if "conv3d_k3_c48_tc" in kernels:
        dominant = dict(kernels["conv3d_k3_c48_tc"])
        dominant["kernel"] = "conv3d_k3_c48_roll_kernel on 2x48x128^3 (conv2 of encoder1 / decoder1, both passes of decoder1.conv1)"
        dominant["share_of_forward"] = "4 launches of the window forward (ncu launch list, profiles/r02_launch_summary.txt)"
        # dram bytes per launch from ncu (profiles/r02_k3_roll_dram.csv): 751 MB read + 381 MB written for 403 + 403 MB algorithmic.  Each
        # input row is needed by three planes; the CTAs walk their first plane in lockstep (neighbours share through L2), the evenly split
        # remainder of the planes still re-reads (with one contiguous run per CTA the input was read three times: 1 212 MB).
        dominant["traffic"] = 1132000000
    extras = {}
    if world == 1 and not args.no_extras:
        if dtype == torch.bfloat16:
            extras["parity"] = parity_probe(dev, eager)
        # 8-pass mirror TTA (light_training/prediction.py:110-160) through the re-hosted Predictor: the flips are index
        # transforms inside the gather / accumulate / normalise kernels, the mean is accumulated by the normalise kernel
        from waveformer_b200.prediction import Predictor
        pred = Predictor(SlidingWindowInferer(**kw), mirror_axes=[0, 1, 2])
        with torch.no_grad():
            pred.maybe_mirror_and_predict_cuda(resident[:1], model)
            ms_tta = event_ms(lambda: pred.maybe_mirror_and_predict_cuda(resident[:1], model), 2, warm=0)
        extras["tta8"] = dict(ms_per_volume=ms_tta, voxels_per_s=VOXELS / (ms_tta * 1e-3), passes=8,
                              vs_single_pass=ms_tta / ms_step * volumes,
                              api="waveformer_b200.prediction.Predictor(inferer, mirror_axes=[0, 1, 2])")
        del resident
        torch.cuda.empty_cache()
        extras["train_step"] = train_step_probe(dev)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        sec, times = cpu_patch_seconds(3, 1)
        cpu = dict(value=VOXELS / (WINDOWS_PER_VOLUME * sec), unit="voxels/s", cores=host_threads(), kind="port",
                   sample=f"1 of 18 windows: one 1x4x128^3 fp32 forward of the oracle port, median of {len(times)} timed forwards "
                          "after 1 warm-up; volume time = 18 x patch time", patch_seconds=[round(t, 4) for t in times])
    line = dict(metric="sliding_window_voxels_per_s", value=value, unit="voxels/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling="weak" if world == 1 else "strong",
                vs_baseline=None, dtype=args.dtype, data="synthetic", config=workload_config(volumes, world, args.sw_batch),
                clocks=clocks, e2e=e2e, gpu_launches=launches, roofline=roof, roofline_dominant_by_time=dominant,
                roofline_kernels=kernels, cpu_baseline=cpu, strong=strong, sw_batch_2=sw2, **extras)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
