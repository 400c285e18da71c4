"""GPU parity of the block-glue kernels (depthwise 3^3 conv, fused InstanceNorm) vs plain PyTorch fp32 on the CPU
(the same functional calls the oracle uses: oracle/model.py ccf_ffn / res_block)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_rel, seeded_randn
from oracle.state import ModelConfig, make_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 8, 8, 8, 192), (1, 5, 7, 9, 48), (1, 4, 4, 6, 20), (1, 16, 16, 16, 384)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 6e-3)])
def test_depthwise_conv_matches_torch(shape, dtype, tol):
    from waveformer_b200 import ops
    B, D, H, W, C = shape
    x = seeded_randn(shape, 40).to(dtype)
    w = seeded_randn((C, 1, 3, 3, 3), 41) * 0.3
    b = seeded_randn((C,), 42) * 0.1
    want = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w, b, padding=1, groups=C).permute(0, 2, 3, 4, 1)
    got = ops.dwconv3d_channels_last(x.cuda(), ops.repack_depthwise_weight(w.cuda()), b.cuda())
    assert got.dtype == dtype and got.shape == x.shape
    assert max_rel(got.float().cpu(), want) < tol
    nob = ops.dwconv3d_channels_last(x.cuda(), ops.repack_depthwise_weight(w.cuda()), None)
    assert max_rel(nob.float().cpu(), want - b) < tol


@pytest.mark.parametrize("shape", [(2, 48, 16, 16, 16), (1, 96, 6, 10, 4), (2, 20, 4, 4, 4), (1, 384, 8, 8, 8)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 8e-3), (torch.float16, 1e-3)])
@pytest.mark.parametrize("channels_last", [True, False])
def test_instance_norm_act_matches_torch(shape, dtype, tol, channels_last):
    from waveformer_b200 import ops
    x = (seeded_randn(shape, 50) * 3 + 1.5).to(dtype)        # non-zero mean: exercises the E[x^2]-E[x]^2 path
    r = seeded_randn(shape, 51).to(dtype)
    fmt = torch.channels_last_3d if channels_last else torch.contiguous_format
    xc, rc = x.cuda().contiguous(memory_format=fmt), r.cuda().contiguous(memory_format=fmt)
    xf, rf = x.float(), r.float()
    cases = [
        (dict(act="leakyrelu", slope=0.01), F.leaky_relu(F.instance_norm(xf), 0.01)),
        (dict(act="relu"), F.relu(F.instance_norm(xf))),
        (dict(act="none"), F.instance_norm(xf)),
        (dict(act="leakyrelu", slope=0.01, res=rc), F.leaky_relu(F.instance_norm(xf) + rf, 0.01)),
        (dict(act="leakyrelu", slope=0.01, res=rc, res_norm=True), F.leaky_relu(F.instance_norm(xf) + F.instance_norm(rf), 0.01)),
    ]
    for kw, want in cases:
        got = ops.instance_norm_act(xc, **kw)
        assert got.shape == x.shape and got.dtype == dtype
        assert max_rel(got.float().cpu(), want) < tol, kw


def test_instance_norm_fp16_block_writes_bf16_slice():
    """fp16 skip block (precision policy): fp16 input + fp16 identity shortcut, bf16 result into a concat slice."""
    from waveformer_b200 import ops
    x = seeded_randn((2, 48, 8, 8, 8), 54).half().cuda().contiguous(memory_format=torch.channels_last_3d)
    r = seeded_randn((2, 48, 8, 8, 8), 55).half().cuda().contiguous(memory_format=torch.channels_last_3d)
    buf = torch.zeros((2, 8, 8, 8, 96), device="cuda", dtype=torch.bfloat16)
    y = ops.instance_norm_act(x, "leakyrelu", 0.01, res=r, out=buf[..., 48:])
    want = F.leaky_relu(F.instance_norm(x.float().cpu()) + r.float().cpu(), 0.01)
    assert y.dtype == torch.bfloat16 and max_rel(y.float().cpu(), want) < 8e-3
    assert bool((buf[..., :48] == 0).all())


def test_instance_norm_large_offset_is_stable():
    from waveformer_b200 import ops
    x = seeded_randn((1, 8, 32, 32, 32), 52) * 0.01 + 100.0   # |mean| / std = 1e4
    got = ops.instance_norm_act(x.cuda(), "none")
    want = F.instance_norm(x.double()).float()
    assert max_rel(got.cpu(), want) < 2e-3


def test_instance_norm_writes_into_concat_slice():
    from waveformer_b200 import ops
    x = seeded_randn((1, 16, 4, 4, 4), 53).cuda().contiguous(memory_format=torch.channels_last_3d)
    buf = torch.full((1, 4, 4, 4, 40), 5.0, device="cuda")
    y = ops.instance_norm_act(x, "relu", out=buf[..., 8:24])
    assert torch.equal(y.permute(0, 2, 3, 4, 1), buf[..., 8:24])
    assert bool((buf[..., :8] == 5).all()) and bool((buf[..., 24:] == 5).all())
    assert max_rel(y.cpu(), F.relu(F.instance_norm(x.cpu()))) < 5e-6


@pytest.mark.parametrize("shape", [(2, 4, 16, 24, 32), (2, 4, 6, 10, 12), (1, 4, 8, 8, 8), (3, 4, 5, 7, 9)])
@pytest.mark.parametrize("in_dtype", [torch.float32, torch.bfloat16])
def test_conv3d_c4_with_fused_shortcut_and_statistics(shape, in_dtype):
    """tcgen05 implicit-GEMM first-block convolution: conv1 (3^3) + conv3 (1^3) + both InstanceNorm statistics vs
    torch's convolutions on the same bf16-rounded operands (ragged tiles, tiles straddling two batch elements)."""
    from waveformer_b200 import ops
    x = seeded_randn(shape, 70).cuda().to(in_dtype).contiguous(memory_format=torch.channels_last_3d)
    w1 = (seeded_randn((48, 4, 3, 3, 3), 71) / 108 ** 0.5).cuda().bfloat16()
    w3 = (seeded_randn((48, 4, 1, 1, 1), 72) / 2.0).cuda().bfloat16()
    y0, s0, y1, s1 = ops.conv3d_c4_in_stats(x, w1, w3, eps=1e-5)
    xr = x.bfloat16().float()
    want0 = F.conv3d(xr, w1.float(), padding=1)
    want1 = F.conv3d(xr, w3.float())
    assert y0.dtype == torch.bfloat16 and tuple(y0.shape) == tuple(want0.shape)
    assert max_rel(y0.float().cpu(), want0.cpu()) < 6e-3
    assert max_rel(y1.float().cpu(), want1.cpu()) < 6e-3
    for y, s in ((y0, s0), (y1, s1)):
        f = y.float()
        mean = f.mean(dim=(2, 3, 4)).reshape(-1)
        rstd = (f.var(dim=(2, 3, 4), unbiased=False) + 1e-5).rsqrt().reshape(-1)
        got = s.reshape(-1, 2)
        assert float((got[:, 0] - mean).abs().max()) < 2e-5 * max(1.0, float(mean.abs().max()))
        assert max_rel(got[:, 1].cpu(), rstd.cpu()) < 1e-4
    # conv only (no shortcut)
    y0b, s0b, y1b, s1b = ops.conv3d_c4_in_stats(x, w1, None)
    assert y1b is None and s1b is None and torch.equal(y0b, y0)


@pytest.mark.parametrize("out_dtype", [torch.float16, torch.bfloat16])
def test_instance_norm_fp32_block_writes_16bit_slice(out_dtype):
    """An fp32 block writing into a 16-bit concatenation slice (8-byte narrow stores): fp16 destinations used to receive bf16 bit
    patterns."""
    from waveformer_b200 import ops
    x = (seeded_randn((2, 16, 6, 10, 12), 58) * 2.0 + 0.5).cuda().contiguous(memory_format=torch.channels_last_3d)
    buf = torch.zeros((2, 6, 10, 12, 40), device="cuda", dtype=out_dtype)
    y = ops.instance_norm_act(x, "leakyrelu", 0.01, out=buf[..., 8:24])
    want = F.leaky_relu(F.instance_norm(x), 0.01)
    assert y.dtype == out_dtype and torch.equal(y.permute(0, 2, 3, 4, 1), buf[..., 8:24])
    assert max_rel(y.float().cpu(), want.cpu()) < (1e-2 if out_dtype == torch.bfloat16 else 2e-3)
    assert bool((buf[..., :8] == 0).all()) and bool((buf[..., 24:] == 0).all())


@pytest.mark.parametrize("shape", [(2, 4, 16, 24, 32), (3, 4, 5, 7, 9), (1, 4, 40, 40, 40)])
@pytest.mark.parametrize("in_dtype,op_dtype", [(torch.float32, torch.float16), (torch.float16, torch.float16),
                                               (torch.bfloat16, torch.bfloat16)])
def test_first_block_shortcut_recomputed_in_the_last_pass(shape, in_dtype, op_dtype):
    """conv3d_c4_in_stats(store_shortcut=False) + instance_norm_act_shortcut4 (the 1^3 shortcut `norm3(conv3(inp))` of
    dynunet_block.py:104-110 recomputed per voxel from the 4-channel input) vs the stored-shortcut path and vs torch."""
    from waveformer_b200 import ops
    x = seeded_randn(shape, 73).cuda().to(in_dtype).contiguous(memory_format=torch.channels_last_3d)
    w1 = (seeded_randn((48, 4, 3, 3, 3), 71) / 108 ** 0.5).cuda().to(op_dtype)
    w3 = (seeded_randn((48, 4, 1, 1, 1), 72) / 2.0).cuda().to(op_dtype)
    y0, s0, y1, s1 = ops.conv3d_c4_in_stats(x, w1, w3, eps=1e-5)
    y0n, s0n, y1n, s1n = ops.conv3d_c4_in_stats(x, w1, w3, eps=1e-5, store_shortcut=False)
    assert y1n is None and torch.equal(y0n, y0) and torch.equal(s0n, s0) and torch.equal(s1n, s1)
    stored = ops.instance_norm_act(y0, "leakyrelu", 0.01, res=y1, res_norm=True, stats=s0, res_stats=s1)
    buf = torch.full(tuple(y0.permute(0, 2, 3, 4, 1).shape[:-1]) + (64,), 3.0, device="cuda", dtype=op_dtype)
    got = ops.instance_norm_act_shortcut4(y0, x, w3, s0, s1, "leakyrelu", 0.01, out=buf[..., 8:56])
    assert torch.equal(got.permute(0, 2, 3, 4, 1), buf[..., 8:56]) and bool((buf[..., :8] == 3).all()) and bool((buf[..., 56:] == 3).all())
    # the shortcut's statistics from the moments of the 4-channel input (no convolution): against those of the unrounded shortcut
    xr = x.to(op_dtype).float()
    c3 = F.conv3d(xr.double(), w3.double())
    sm = ops.shortcut4_stats(x, w3, eps=1e-5).reshape(-1, 2)
    mean = c3.mean(dim=(2, 3, 4)).reshape(-1)
    rstd = (c3.var(dim=(2, 3, 4), unbiased=False) + 1e-5).rsqrt().reshape(-1)
    assert float((sm[:, 0].double() - mean).abs().max()) < 2e-6 * max(1.0, float(mean.abs().max()))
    assert max_rel(sm[:, 1].cpu(), rstd.float().cpu()) < 1e-5
    got_m = ops.instance_norm_act_shortcut4(y0, x, w3, s0, sm.reshape(-1), "leakyrelu", 0.01)
    want = F.leaky_relu(F.instance_norm(y0.float()) + F.instance_norm(F.conv3d(xr, w3.float())), 0.01)
    assert max_rel(got_m.float().cpu(), want.cpu()) < (1e-2 if op_dtype == torch.bfloat16 else 2e-3)
    tol = 1e-2 if op_dtype == torch.bfloat16 else 2e-3
    assert max_rel(got.float().cpu(), want.cpu()) < tol
    assert max_rel(got.float().cpu(), stored.float().cpu()) < 2 * tol      # the stored shortcut is rounded to 16 bit first


@pytest.mark.parametrize("shape,dtype", [((2, 48, 20, 24, 28), torch.float16), ((1, 96, 9, 11, 13), torch.bfloat16),
                                         ((3, 144, 5, 6, 7), torch.float16), ((2, 16, 33, 8, 8), torch.float32),
                                         ((1, 384, 4, 4, 4), torch.float16)])
@pytest.mark.parametrize("mode", ["plain", "res", "res_norm", "affine"])
def test_instance_norm_act_register_constant_kernel(shape, dtype, mode):
    """The register-constant, software-pipelined InstanceNorm pass (same-width packets) on ragged sizes: every residual mode
    and the GroupNorm affine, against torch in fp32."""
    from waveformer_b200 import ops
    x = (seeded_randn(shape, 54) * 1.3 + 0.4).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    r = seeded_randn(shape, 55).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    c = shape[1]
    g = (1.0 + 0.1 * seeded_randn((c,), 56)).cuda()
    b = (0.05 * seeded_randn((c,), 57)).cuda()
    want = F.instance_norm(x.float())
    if mode == "plain":
        got = ops.instance_norm_act(x, "leakyrelu", 0.01)
    elif mode == "res":
        got, want = ops.instance_norm_act(x, "leakyrelu", 0.01, res=r), want + r.float()
    elif mode == "res_norm":
        got, want = ops.instance_norm_act(x, "leakyrelu", 0.01, res=r, res_norm=True), want + F.instance_norm(r.float())
    else:
        got, want = ops.instance_norm_act(x, "leakyrelu", 0.01, gamma=g, beta=b), want * g.view(1, -1, 1, 1, 1) + b.view(1, -1, 1, 1, 1)
    want = F.leaky_relu(want, 0.01)
    tol = 5e-6 if dtype == torch.float32 else (1e-2 if dtype == torch.bfloat16 else 2e-3)
    assert got.dtype == dtype and max_rel(got.float().cpu(), want.cpu()) < tol


@pytest.mark.parametrize("rows,c", [(1000, 48), (513, 96), (300, 192), (77, 384), (40, 768), (9, 1536), (64, 20), (50, 8)])
@pytest.mark.parametrize("in_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("gelu", [False, True])
def test_layer_norm_channels_last_matches_torch(rows, c, in_dtype, out_dtype, gelu):
    """LayerNorm (+ GELU erf) over channels: every lane-per-row configuration of the wide kernel, the narrow fallback
    (C % 8 != 0), mixed storage types and the dual fp32 + bf16 output."""
    from waveformer_b200 import ops
    x = (seeded_randn((rows, c), 80) * 1.7 + 0.3).cuda().to(in_dtype)
    g = (1.0 + 0.1 * seeded_randn((c,), 81)).cuda()
    b = (0.05 * seeded_randn((c,), 82)).cuda()
    want = F.layer_norm(x.float(), (c,), g, b, 1e-6)
    if gelu:
        want = F.gelu(want)
    got = ops.layer_norm_cl(x, g, b, 1e-6, gelu=gelu, out_dtype=out_dtype)
    tol = 2e-6 if out_dtype == torch.float32 and in_dtype == torch.float32 else 5e-3
    assert got.dtype == out_dtype and max_rel(got.float().cpu(), want.cpu()) < tol
    if out_dtype == torch.float32:
        y, y2 = ops.layer_norm_cl(x, g, b, 1e-6, gelu=gelu, out_dtype=out_dtype, also_bf16=True)
        assert torch.equal(y, got) and torch.equal(y2, got.bfloat16())
    # affine-free (proj_out)
    got0 = ops.layer_norm_cl(x, None, None, 1e-5, out_dtype=out_dtype)
    assert max_rel(got0.float().cpu(), F.layer_norm(x.float(), (c,)).cpu()) < tol


@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-6), (torch.bfloat16, 8e-3)])
def test_trilinear_upsample_sum_matches_torch(align, dtype, tol):
    """base + sum of trilinear upsamples of 1..3 coarse maps (Block.multi_scale_forward / ProjectionUpsample)."""
    from waveformer_b200 import ops
    size = (16, 24, 32)
    srcs = [seeded_randn((2, 8, 12, 16, 48), 90), seeded_randn((2, 4, 6, 8, 48), 91), seeded_randn((2, 2, 3, 4, 48), 92)]
    base = seeded_randn((2,) + size + (48,), 93)
    for n in (1, 3):
        want = base.clone() if not align else torch.zeros_like(base)
        for s in srcs[:n]:
            s_r = s.to(dtype).float()
            want = want + F.interpolate(s_r.permute(0, 4, 1, 2, 3), size=size, mode="trilinear",
                                        align_corners=align).permute(0, 2, 3, 4, 1)
        got = ops.upsample_trilinear_add([s.cuda().to(dtype) for s in srcs[:n]], size,
                                         base=None if align else base.cuda(), align_corners=align,
                                         out_dtype=torch.float32 if not align else dtype)
        assert max_rel(got.float().cpu(), want) < tol


@pytest.mark.parametrize("shape", [(2, 16, 16, 32, 192), (1, 9, 10, 23, 96), (1, 8, 8, 8, 1536), (1, 4, 6, 40, 72)])
def test_depthwise_conv_tile_kernel_matches_register_kernel(shape):
    """The shared-memory FHFMA kernel (bf16) against torch and against the register-tiled kernel (WF_DWCONV_IMPL=reg):
    ragged tiles in every axis, partial channel groups (C % 64 != 0), more than one channel group."""
    import os
    from waveformer_b200 import ops
    c = shape[-1]
    x = seeded_randn(shape, 95).cuda().bfloat16()
    w = (seeded_randn((c, 1, 3, 3, 3), 96) / 27 ** 0.5)
    bias = (0.05 * seeded_randn((c,), 97)).cuda()
    w27 = ops.repack_depthwise_weight(w.cuda())
    got = ops.dwconv3d_channels_last(x, w27, bias)
    os.environ["WF_DWCONV_IMPL"] = "reg"
    try:
        old = ops.dwconv3d_channels_last(x, w27, bias)
    finally:
        os.environ.pop("WF_DWCONV_IMPL", None)
    want = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.cuda().bfloat16().float(), bias, padding=1, groups=c).permute(0, 2, 3, 4, 1)
    assert max_rel(got.float().cpu(), want.cpu()) < 6e-3          # bf16 taps, fp32 accumulation, bf16 result
    assert max_rel(got.float().cpu(), old.float().cpu()) < 8e-3   # register kernel keeps fp32 taps


@pytest.mark.parametrize("shape,k", [((2, 48, 8, 12, 16), 4), ((1, 96, 5, 6, 7), 13), ((3, 16, 4, 4, 4), 2)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("res_norm", [True, False])
def test_instance_norm_act_with_fused_output_head(shape, k, dtype, tol, res_norm):
    """act(IN(x) + IN(res) | res) -> 1^3 conv (+ bias) in one kernel vs the two-step torch computation."""
    from waveformer_b200 import ops
    x = seeded_randn(shape, 110).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    r = seeded_randn(shape, 111).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = (seeded_randn((k, shape[1], 1, 1, 1), 112) / shape[1] ** 0.5).cuda()
    b = (0.1 * seeded_randn((k,), 113)).cuda()
    got = ops.instance_norm_act_head(x, w, b, "leakyrelu", 0.01, res=r, res_norm=res_norm, out_dtype=torch.float32)
    xc, rc = x.double().cpu(), r.double().cpu()                     # reference on the host in fp64 (no TF32 anywhere)
    t = F.instance_norm(xc) + (F.instance_norm(rc) if res_norm else rc)
    want = F.conv3d(F.leaky_relu(t, 0.01), w.double().cpu(), b.double().cpu())
    assert got.dtype == torch.float32 and tuple(got.shape) == tuple(want.shape)
    assert max_rel(got.cpu(), want) < tol


@pytest.mark.parametrize("shape,cout", [((2, 8, 8, 16, 144), 48), ((1, 3, 5, 7, 32), 16), ((1, 4, 4, 4, 64), 64)])
def test_conv_transpose_k2s2_scatters_into_concat_buffer(shape, cout):
    """ConvTranspose3d(k=2, s=2) as one tcgen05 GEMM writing channels [0, Cout) of a wider channels-last buffer."""
    from waveformer_b200 import ops
    cin = shape[-1]
    x = seeded_randn(shape, 120).cuda().bfloat16()
    w = (seeded_randn((cin, cout, 2, 2, 2), 121) / cin ** 0.5).cuda().bfloat16()
    B, D, H, W, _ = shape
    cat = torch.full((B, 2 * D, 2 * H, 2 * W, 2 * cout), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.conv_transpose3d_k2s2(x, w, out=cat[..., :cout])
    want = F.conv_transpose3d(x.float().permute(0, 4, 1, 2, 3), w.float(), stride=2).permute(0, 2, 3, 4, 1)
    assert max_rel(cat[..., :cout].float().cpu(), want.cpu()) < 6e-3
    assert bool((cat[..., cout:] == 7.0).all())                      # the skip half is untouched


def test_trilinear_upsample_generic_ratio_uses_point_kernel():
    """Scale factors below 2 fall back to the one-output-per-thread kernel (the 2x2x2-cell kernel needs >= 2)."""
    from waveformer_b200 import ops
    src = seeded_randn((1, 8, 12, 16, 24), 130)
    size = (12, 18, 24)
    want = F.interpolate(src.permute(0, 4, 1, 2, 3), size=size, mode="trilinear", align_corners=False).permute(0, 2, 3, 4, 1)
    got = ops.upsample_trilinear_add([src.cuda()], size)
    assert max_rel(got.cpu(), want) < 3e-6


@pytest.mark.parametrize("c_dtype", [torch.float32, torch.bfloat16])
def test_residual_sum_matches_torch(c_dtype):
    from waveformer_b200 import ops
    a, b = seeded_randn((3, 5, 7, 9, 48), 140).cuda(), seeded_randn((3, 5, 7, 9, 48), 141).cuda()
    c = seeded_randn((3, 5, 7, 9, 48), 142).cuda().to(c_dtype)
    bias = seeded_randn((48,), 143).cuda()
    got = ops.residual_sum(a, b, c, bias)
    want = a + b + c.float() + bias
    assert got.dtype == torch.float32 and float((got - want).abs().max()) < 1e-6
    assert float((ops.residual_sum(a, b, c) - (a + b + c.float())).abs().max()) < 1e-6


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 5e-3)])
def test_gelu_inplace_matches_torch(dtype, tol):
    from waveformer_b200 import ops
    x = (seeded_randn((3, 7, 16, 8), 150) * 2.5).cuda().to(dtype)
    want = F.gelu(x.float())
    got = ops.gelu_(x.clone())
    assert got.dtype == dtype and max_rel(got.float().cpu(), want.cpu()) < tol


@pytest.mark.parametrize("shape", [(1, 2, 4, 128), (2, 3, 6, 128), (1, 5, 9, 128), (1, 1, 1, 128), (2, 37, 11, 128), (2, 24, 40, 128), (2, 128, 128, 128)])
@pytest.mark.parametrize("fused_input_norm", [False, True])
def test_conv3d_k3_c48_producer_consumer_kernel(shape, fused_input_norm):
    """tcgen05 3^3 convolution 48 -> 48 on rows of 128 voxels (ring-staged input rows, shifted-descriptor dx taps, double-
    buffered TMEM accumulators): vs torch's convolution on the same bf16-rounded operands; fused InstanceNorm + LeakyReLU of
    the input; fused output statistics; partial row blocks (H % 4 != 0) and volume borders in z / y / x.  The last two
    shapes make every persistent CTA walk several / 56 row blocks (ring wrap-around, both TMEM accumulator buffers many
    times over); the last one is BASELINE's full size."""
    from waveformer_b200 import ops
    B, D, H, W = shape
    x = (seeded_randn((B, 48, D, H, W), 160) * 1.3 + 0.2).cuda().bfloat16().contiguous(memory_format=torch.channels_last_3d)
    w = (seeded_randn((48, 48, 3, 3, 3), 161) / (27 * 48) ** 0.5).cuda().bfloat16()
    xin = x.float()
    stats = None
    if fused_input_norm:
        stats = ops.instance_norm_stats(x, eps=1e-5)
        t = F.leaky_relu(F.instance_norm(xin, eps=1e-5), 0.01)
        xin = t.bfloat16().float()                      # the kernel rounds the normalised operand to bf16 as well
    want = F.conv3d(xin, w.float(), padding=1)
    y, st = ops.conv3d_k3_c48(x, w, in_stats=stats, slope=0.01)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == tuple(want.shape)
    assert max_rel(y.float().cpu(), want.cpu()) < 8e-3
    f = y.float()
    mean = f.mean(dim=(2, 3, 4)).reshape(-1)
    rstd = (f.var(dim=(2, 3, 4), unbiased=False) + 1e-5).rsqrt().reshape(-1)
    got = st.reshape(-1, 2)
    assert float((got[:, 0] - mean).abs().max()) < 2e-5 * max(1.0, float(mean.abs().max()))
    assert max_rel(got[:, 1].cpu(), rstd.cpu()) < 1e-4


@pytest.mark.parametrize("shape", [(1, 2, 4, 128), (2, 3, 7, 128), (1, 1, 1, 128), (2, 37, 11, 128), (2, 128, 128, 128)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_conv3d_k3_c96_c48_two_pass(shape, dtype):
    """decoder1's first convolution (3^3, 96 -> 48 on the concatenation buffer) as two passes of the rolling-row 48 -> 48 kernel,
    the second adding the first's result to its fp32 accumulators: vs torch's convolution of the same 16-bit operands (the
    partial sum is rounded to 16 bits once: tolerance 2 roundings).  Odd D / H make the per-CTA runs of output rows start and
    end mid-plane and cross planes and batch elements; the last shape is BASELINE's full size."""
    from waveformer_b200 import ops
    B, D, H, W = shape
    x = (seeded_randn((B, 96, D, H, W), 170) * 1.1 - 0.1).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = (seeded_randn((48, 96, 3, 3, 3), 171) / (27 * 96) ** 0.5).cuda().to(dtype)
    want = F.conv3d(x.float(), w.float(), padding=1)
    y, st = ops.conv3d_k3_c96_c48(x, w)
    assert y.dtype == dtype and tuple(y.shape) == tuple(want.shape)
    assert max_rel(y.float().cpu(), want.cpu()) < (2e-3 if dtype == torch.float16 else 1.2e-2)
    f = y.float()
    mean = f.mean(dim=(2, 3, 4)).reshape(-1)
    rstd = (f.var(dim=(2, 3, 4), unbiased=False) + 1e-5).rsqrt().reshape(-1)
    got = st.reshape(-1, 2)
    assert float((got[:, 0] - mean).abs().max()) < 2e-5 * max(1.0, float(mean.abs().max()))
    assert max_rel(got[:, 1].cpu(), rstd.cpu()) < 1e-4


@pytest.mark.parametrize("shape", [(2, 8, 12, 16, 48), (1, 4, 4, 8, 96), (1, 4, 4, 4, 192), (2, 2, 6, 4, 16)])
@pytest.mark.parametrize("out_dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 8e-3)])
@pytest.mark.parametrize("v2", [False, True])
def test_patch_merge_gather_layernorm_fused(shape, out_dtype, tol, v2):
    """Octant gather + LayerNorm(8C) in one kernel vs torch.cat + F.layer_norm, for MONAI 0.9's order (two octants
    repeated, reference wave_helper.py:170-194) and the itertools.product order of PatchMergingV2."""
    import itertools
    from waveformer_b200 import ops
    monai09 = ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1), (1, 1, 1))
    octs = tuple(itertools.product(range(2), range(2), range(2))) if v2 else monai09
    x = seeded_randn(shape, 140) * 2 + 0.5
    c8 = 8 * shape[-1]
    gam, bet = 1 + 0.1 * seeded_randn((c8,), 141), 0.1 * seeded_randn((c8,), 142)
    cat = torch.cat([x[:, i::2, j::2, k::2, :] for i, j, k in octs], -1)
    want = F.layer_norm(cat, (c8,), gam, bet, 1e-5)
    got = ops.patch_merge_layer_norm(x.cuda(), gam.cuda(), bet.cuda(), 1e-5, octs, out_dtype)
    assert got is not None and got.dtype == out_dtype and tuple(got.shape) == tuple(want.shape)
    assert max_rel(got.float().cpu(), want) < tol
    assert ops.patch_merge_layer_norm(x[:, :, :, :-1].contiguous().cuda(), gam.cuda(), bet.cuda(), 1e-5, octs, out_dtype) is None


@pytest.mark.parametrize("k2,rows", [(0, 128), (0, 777), (96, 128), (96, 1000), (192, 257), (192, 4 * 32 * 32 * 32)])
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 3e-3), (torch.bfloat16, 2e-2)])
def test_pw_gelu_dual_tail_kernel(k2, rows, dtype, tol):
    """Tail of ProjectionUpsample in one tcgen05 kernel: out = W1 . GELU(h) + b1 + W2 . u + b2 written into a channel slice of a
    wider channels-last buffer (row pitch 144), vs fp64 torch on the same 16-bit operands; partial last tile, several tiles per CTA."""
    from waveformer_b200 import ops
    h = (seeded_randn((rows, 192), 180) * 1.5).cuda().to(dtype)
    w1 = (seeded_randn((48, 192, 1, 1, 1), 182) / 192 ** 0.5).cuda()
    b1, b2 = (0.1 * seeded_randn((48,), 184)).cuda(), (0.1 * seeded_randn((48,), 185)).cuda()
    comb = torch.full((rows, 144), 7.0, device="cuda", dtype=dtype)
    out = comb[:, 48:96]
    w1h = w1.view(48, 192).to(dtype).double()
    want = F.gelu(h.double()) .to(dtype).double() @ w1h.t() + b1.double()
    if k2:      # residual as a second product on the upsampled input
        u = seeded_randn((rows, k2), 181).cuda().to(dtype)
        w2 = (seeded_randn((48, k2, 1, 1, 1), 183) / k2 ** 0.5).cuda()
        assert ops.pw_gelu_dual_supported(h, u, 48, out)
        ops.pw_gelu_dual(h, u, w1, b1, w2, b2, out)
        want = want + u.double() @ w2.view(48, k2).to(dtype).double().t() + b2.double()
    else:       # residual as an fp32 addend (the commuted low-resolution convolution, upsampled)
        add = seeded_randn((rows, 48), 186).cuda()
        assert ops.pw_gelu_dual_supported(h, None, 48, out)
        ops.pw_gelu_dual(h, None, w1, b1, None, None, out, addend=add)
        want = want + add.double()
    assert max_rel(out.double().cpu(), want.cpu()) < tol
    assert bool((comb[:, :48] == 7.0).all()) and bool((comb[:, 96:] == 7.0).all())      # neighbouring slices untouched


@pytest.mark.parametrize("cin,cout,stride,double", [(96, 48, 2, False), (192, 48, 4, True), (16, 8, 2, False)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
def test_projection_upsample_fused_matches_module_math(cin, cout, stride, double, dtype, tol):
    """ProjectionUpsample inference path (cell upsample, depthwise stencil, GroupNorm folded into conv2 by
    wf_groupnorm_fold_linear, GELU kernel, last conv + residual projection accumulated into a concat slice) vs the
    module's own torch ops (= the reference's wave_helper.py:33-81) in fp64 on the host."""
    from waveformer_b200.network_models.wave_helper import ProjectionUpsample
    torch.manual_seed(5)
    m = ProjectionUpsample(cin, cout, stride=stride, residual=True, use_double_conv=double).eval()
    with torch.no_grad():
        m.norm.weight.copy_(1 + 0.2 * seeded_randn((cin,), 150))
        m.norm.bias.copy_(0.2 * seeded_randn((cin,), 151))
    x = seeded_randn((2, cin, 4, 4, 4), 152)
    with torch.no_grad():
        want = m.double()(x.double()).float()
    m = m.float().cuda()
    if dtype == torch.bfloat16:
        keep = {id(p) for p in list(m.norm.parameters()) + list(m.conv1[1].parameters())}
        for p in m.parameters():
            if id(p) not in keep:
                p.data = p.data.bfloat16()
    xc = x.cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    buf = torch.zeros((2,) + tuple(want.shape[2:]) + (cout + 8,), device="cuda", dtype=dtype)
    with torch.no_grad():
        got = m(xc, out_buf=buf[..., 8:])
    assert tuple(got.shape) == tuple(want.shape)
    assert max_rel(got.float().cpu(), want) < tol
    assert bool((buf[..., :8] == 0).all())
    gamma, beta = m.norm.weight, m.norm.bias
    mr = torch.stack([seeded_randn((2 * cin,), 153), seeded_randn((2 * cin,), 154).abs() + 0.5], -1).reshape(-1).cuda()
    from waveformer_b200 import ops
    w = seeded_randn((2 * cin, cin), 155).cuda()
    b = seeded_randn((2 * cin,), 156).cuda()
    wf_, bf_ = ops.groupnorm_fold_linear(mr, gamma, beta, w, b, 2, torch.float32)
    a = mr.view(2, cin, 2)[..., 1] * gamma
    d = beta - mr.view(2, cin, 2)[..., 0] * a
    assert max_rel(wf_.cpu(), (w[None] * a[:, None, :]).cpu()) < 1e-6
    assert max_rel(bf_.cpu(), (b[None] + d @ w.t()).cpu()) < 1e-5


@pytest.mark.parametrize("shape", [(2, 8, 8, 16, 192), (1, 5, 7, 9, 48), (2, 4, 4, 8, 24)])
def test_depthwise_conv_with_fused_statistics(shape):
    """bf16 tile kernel with the following normalisation's statistics reduced in its epilogue: same output as the plain
    kernel, (mean, rstd) equal to the separate statistics pass over that output (ragged tiles, partial channel groups)."""
    from waveformer_b200 import ops
    C = shape[-1]
    x = (seeded_randn(shape, 160) + 0.3).bfloat16().cuda()
    w = ops.repack_depthwise_weight((seeded_randn((C, 1, 3, 3, 3), 161) * 0.3).cuda())
    b = (seeded_randn((C,), 162) * 0.1).cuda()
    y, mr = ops.dwconv3d_channels_last_stats(x, w, b, 1e-5)
    want_y = ops.dwconv3d_channels_last(x, w, b)
    assert torch.equal(y, want_y)
    want = ops.instance_norm_stats(want_y.permute(0, 4, 1, 2, 3), eps=1e-5)
    assert max_rel(mr.view(-1, 2)[:, 0].cpu(), want.view(-1, 2)[:, 0].cpu()) < 1e-4
    assert max_rel(mr.view(-1, 2)[:, 1].cpu(), want.view(-1, 2)[:, 1].cpu()) < 1e-4


@pytest.mark.parametrize("C,rows", [(48, 128 * 37 + 5), (96, 128 * 9 + 77), (48, 64)])
@pytest.mark.parametrize("fmt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 1.5e-2)])
def test_fused_ffn_kernels_match_module_math(C, rows, fmt, tol):
    """wf_ffn_front / wf_ffn_back (CCF_FFN's pointwise GEMMs fused with norm2, both LayerNorm + GELU stages, the fc bias and
    both residuals; reference wave_helper.py:260-294,509) vs the same arithmetic in fp32 torch ops."""
    import torch.nn as nn
    import torch.nn.functional as F
    from waveformer_b200 import ops
    g = torch.Generator().manual_seed(C + rows)
    x = (torch.randn((rows, C), generator=g) * 1.5 + 0.3).cuda()
    norm2 = nn.LayerNorm(C, eps=1e-6).cuda()
    ln1, ln2 = nn.LayerNorm(4 * C).cuda(), nn.LayerNorm(4 * C).cuda()
    for ln in (norm2, ln1, ln2):
        ln.weight.data = (1.0 + 0.2 * torch.randn(ln.weight.shape, generator=g)).cuda()
        ln.bias.data = (0.1 * torch.randn(ln.bias.shape, generator=g)).cuda()
    w1 = (torch.randn((4 * C, C), generator=g) / C ** 0.5).cuda()
    b1 = (0.2 * torch.randn((4 * C,), generator=g)).cuda()
    wfc = (torch.randn((C, 4 * C), generator=g) / (4 * C) ** 0.5).cuda()
    bfc = (0.2 * torch.randn((C,), generator=g)).cuda()
    with torch.no_grad():
        n = norm2(x)
        want1 = F.gelu(ln1(F.linear(n, w1, b1)))
        got1 = ops.ffn_front(x, norm2, w1, b1, ln1, fmt)
        assert got1.dtype == fmt and tuple(got1.shape) == (rows, 4 * C)
        assert max_rel(got1.float(), want1) < tol
        t2 = (torch.randn((rows, 4 * C), generator=g) * 0.7 + 0.1).cuda().to(fmt)
        want2 = x + n + F.linear(F.gelu(ln2(t2.float())), wfc, bfc)
        got2 = ops.ffn_back(t2, ln2, wfc, bfc, x, norm2)
        assert got2.dtype == torch.float32 and max_rel(got2, want2) < tol


def test_block_with_fused_ffn_equals_unfused_path(monkeypatch):
    """A stage-1 block under the 16-bit policy: the fused FFN kernels and the unfused sequence of launches agree."""
    from waveformer_b200 import prepare_inference
    from waveformer_b200.network_models import Waveformer
    cfg = ModelConfig(img_size=(64,) * 3)
    m = Waveformer(**cfg.kwargs()).eval()
    m.load_state_dict(make_state_dict(cfg, seed=0), strict=True)
    m = prepare_inference(m.cuda(), torch.bfloat16)
    blk = m.waveformer_encoder.block1[0]
    x = seeded_randn((2, 32, 32, 32, 48), 123).cuda()
    with torch.no_grad():
        monkeypatch.setenv("WF_FFN_FUSED", "0")
        want, _ = blk(x)
        monkeypatch.setenv("WF_FFN_FUSED", "1")
        got, _ = blk(x)
    assert max_rel(got, want) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape,cout", [((2, 4, 16, 24, 32), 48), ((1, 4, 6, 10, 14), 96)])
def test_patch_embed_kernel_matches_conv3d(shape, cout, dtype):
    """wf_patch_embed_k2s2_c4 (PatchEmbed.proj + the rearrange, patchembedding.py:196-225 / waveformer.py:260-270) vs the
    library convolution in fp32 without TF32: exact-fp32 arithmetic, channels-last stream out."""
    from waveformer_b200 import ops
    g = torch.Generator().manual_seed(cout)
    x = torch.randn(shape, generator=g).cuda().to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = (torch.randn((cout, 4, 2, 2, 2), generator=g) * 0.3).cuda()
    b = (torch.randn((cout,), generator=g) * 0.1).cuda()
    got = ops.patch_embed_k2s2(x, w, b)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        want = F.conv3d(x.float(), w, b, stride=2).permute(0, 2, 3, 4, 1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert got.dtype == torch.float32 and got.is_contiguous() and tuple(got.shape) == tuple(want.shape)
    assert max_rel(got, want) < 2e-6
    assert ops.patch_embed_k2s2(x[:, :3], w[:, :3], b) is None          # other channel counts keep the library path
