#!/usr/bin/env python
"""Run ONE hand-written kernel on its production geometry a few times (the command `ncu --set full` wraps).

    python scripts/kernel_cases.py --case haar|c4|dwconv|attn|layernorm|convt|head [--iters 3]
Prints the CUDA-event time per call (a number printed under ncu is not a bench value).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveformer_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", required=True)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
g = torch.Generator("cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731

if args.case == "haar":      # BASELINE configs[1]
    x = rn(2, 48, 128, 128, 128).bfloat16()
    ll, hf = ops._dwt_ncdhw_raw(x, True)
    fns = {"dwt3d_ncdhw_bf16": lambda: ops._dwt_ncdhw_raw(x, True), "idwt3d_ncdhw_bf16": lambda: ops._idwt_ncdhw_raw(ll, hf)}
elif args.case == "c4":      # encoder1: 2 x 4 x 128^3 fp32 window -> conv1 + conv3 + statistics
    x = rn(2, 4, 128, 128, 128).bfloat16().contiguous(memory_format=torch.channels_last_3d)
    w1, w3 = (rn(48, 4, 3, 3, 3) * 0.1).bfloat16(), (rn(48, 4, 1, 1, 1) * 0.5).bfloat16()
    fns = {"conv3d_c4_in_stats": lambda: ops.conv3d_c4_in_stats(x, w1, w3)}
elif args.case == "k3":      # encoder1 / decoder1 conv2: 48 -> 48 at 2 x 128^3, fused input IN + lrelu and output statistics
    x = rn(2, 128, 128, 128, 48).half().permute(0, 4, 1, 2, 3)          # fp16: the 16-bit policy's storage format
    w = (rn(48, 48, 3, 3, 3) / 36).half()
    st = ops.instance_norm_stats(x)
    fns = {"conv3d_k3_c48 (fused in-norm + stats)": lambda: ops.conv3d_k3_c48(x, w, in_stats=st),
           "conv3d_k3_c48 (plain + stats)": lambda: ops.conv3d_k3_c48(x, w),
           "cudnn conv3d 48->48 (library, for comparison)": lambda: torch.nn.functional.conv3d(x, w, padding=1)}
    x96 = rn(2, 128, 128, 128, 96).half().permute(0, 4, 1, 2, 3)
    w96 = (rn(48, 96, 3, 3, 3) / 50).half()
    fns["conv3d_k3_c96_c48 (two passes + stats)"] = lambda: ops.conv3d_k3_c96_c48(x96, w96)
    fns["cudnn conv3d 96->48 (library, for comparison)"] = lambda: torch.nn.functional.conv3d(x96, w96, padding=1)
elif args.case == "instnorm":   # the InstanceNorm passes of the 128^3 residual blocks (2 windows, fp16)
    x = rn(2, 128, 128, 128, 48).half().permute(0, 4, 1, 2, 3)
    r = rn(2, 128, 128, 128, 48).half().permute(0, 4, 1, 2, 3)
    st, rst = ops.instance_norm_stats(x), ops.instance_norm_stats(r)
    x96 = rn(2, 64, 64, 64, 96).half().permute(0, 4, 1, 2, 3)
    st96 = ops.instance_norm_stats(x96)
    fns = {"instance_norm_stats 2x48x128^3 (403 MB read)": lambda: ops.instance_norm_stats(x),
           "instance_norm_act lrelu 2x48x128^3 (403 MB in + 403 MB out)": lambda: ops.instance_norm_act(x, "leakyrelu", 0.01, stats=st),
           "instance_norm_act + IN(res) + lrelu 2x48x128^3 (805 MB in + 403 MB out)":
               lambda: ops.instance_norm_act(x, "leakyrelu", 0.01, res=r, res_norm=True, stats=st, res_stats=rst),
           "instance_norm_act lrelu 2x96x64^3 (101 MB in + 101 MB out)": lambda: ops.instance_norm_act(x96, "leakyrelu", 0.01, stats=st96)}
elif args.case == "dwconv":  # CCF_FFN stage 1: 2 x 64^3 x 192
    x = rn(2, 64, 64, 64, 192).bfloat16()
    w27, b = rn(27, 192) * 0.2, rn(192) * 0.05
    fns = {"dwconv3d_bf16_tile": lambda: ops.dwconv3d_channels_last(x, w27, b)}
elif args.case == "layernorm":
    x = rn(2 * 64 ** 3, 192).bfloat16()
    gam, bet = 1 + 0.1 * rn(192), 0.1 * rn(192)
    fns = {"layernorm_gelu_192_bf16": lambda: ops.layer_norm_cl(x, gam, bet, 1e-5, gelu=True)}
elif args.case == "convt":   # decoder1.transp_conv: 144 -> 48, 64^3 -> 128^3 into the 96-channel concat buffer
    x = rn(2, 64, 64, 64, 144).bfloat16()
    w = (rn(144, 48, 2, 2, 2) / 12).bfloat16()
    cat = torch.empty(2, 128, 128, 128, 96, device="cuda", dtype=torch.bfloat16)
    fns = {"convtranspose_k2s2": lambda: ops.conv_transpose3d_k2s2(x, w, out=cat[..., :48])}
elif args.case == "head":    # decoder1's last kernel: IN + IN(res) + lrelu + 1x1 head, 2 x 128^3 x 48 -> 4
    x = rn(2, 128, 128, 128, 48).bfloat16().permute(0, 4, 1, 2, 3)
    r = rn(2, 128, 128, 128, 48).bfloat16().permute(0, 4, 1, 2, 3)
    w, b = rn(4, 48, 1, 1, 1) / 7, rn(4) * 0.1
    fns = {"instnorm_apply_head": lambda: ops.instance_norm_act_head(x, w, b, "leakyrelu", 0.01, res=r, res_norm=True)}
elif args.case == "attn":    # stage-1 level-1 attention at sw_batch 2
    from waveformer_b200.network_models import Attention
    att = Attention(48, num_heads=3, qkv_bias=True, window_size=8).cuda().eval()
    att.compute_dtype, att.out_dtype, att.split_operands = torch.float16, torch.float32, True
    att2 = Attention(48, num_heads=3, qkv_bias=True, window_size=8).cuda().eval()
    att2.compute_dtype, att2.out_dtype = torch.float16, torch.float32
    x = rn(2, 32, 32, 32, 48)
    fns = {"window_attention_s1L1 (fp16x2: compensated operands, the policy's default)": lambda: att.forward_grid(x),
           "window_attention_s1L1 (plain fp16 operands)": lambda: att2.forward_grid(x)}
elif args.case == "ffn":     # CCF_FFN stage 1 (2 x 64^3 x 48 -> 192 -> 48) and stage 2, fused front / back kernels vs the launches they replace
    import torch.nn as nn
    fns = {}
    for C, n in ((48, 2 * 64 ** 3), (96, 2 * 32 ** 3)):
        x = rn(n, C)
        norm2, ln = nn.LayerNorm(C, eps=1e-6).cuda(), nn.LayerNorm(4 * C).cuda()
        w1, b1, wfc, bfc = rn(4 * C, C) / C ** 0.5, rn(4 * C) * 0.1, rn(C, 4 * C) / (4 * C) ** 0.5, rn(C) * 0.1
        t2 = rn(n, 4 * C).half()
        w1h, wfch = w1.half(), wfc.half()
        fns[f"ffn_front C={C}"] = lambda x=x, norm2=norm2, w1=w1, b1=b1, ln=ln: ops.ffn_front(x, norm2, w1, b1, ln, torch.float16)
        fns[f"ffn_back C={C}"] = lambda t2=t2, ln=ln, wfc=wfc, bfc=bfc, x=x, norm2=norm2: ops.ffn_back(t2, ln, wfc, bfc, x, norm2)

        def unfused_front(x=x, norm2=norm2, w1h=w1h, b1=b1, ln=ln):
            n_, nop = ops.layer_norm_cl(x, norm2.weight, norm2.bias, norm2.eps, also_bf16=torch.float16)
            t = torch.nn.functional.linear(nop, w1h, b1.half())
            return ops.layer_norm_cl(t, ln.weight, ln.bias, ln.eps, gelu=True)

        def unfused_back(t2=t2, ln=ln, wfch=wfch, bfc=bfc, x=x):
            t = ops.layer_norm_cl(t2, ln.weight, ln.bias, ln.eps, gelu=True)
            f = torch.mm(t, wfch.t(), out_dtype=torch.float32)
            return ops.residual_sum(x, x, f, bfc)

        fns[f"unfused front (LN + GEMM + LN/GELU) C={C}"] = unfused_front
        fns[f"unfused back (LN/GELU + GEMM + residual) C={C}"] = unfused_back
else:
    raise SystemExit("unknown case")

with torch.no_grad():
    for name, fn in fns.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        # replayed as a CUDA graph: device time only (an eager loop of ~0.3 ms kernels can be bound by the host's launch path)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(args.iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b_.record()
        torch.cuda.synchronize()
        print(f"{name}: {a.elapsed_time(b_) / args.iters * 1e3:.1f} us per call")
