"""GPU parity of the block-glue kernels (depthwise 3^3 conv, fused InstanceNorm) vs plain PyTorch fp32 on the CPU
(the same functional calls the oracle uses: oracle/model.py ccf_ffn / res_block)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import max_rel, seeded_randn

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
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-6), (torch.bfloat16, 8e-3)])
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
