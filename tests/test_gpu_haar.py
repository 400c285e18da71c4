"""GPU parity: Haar analysis / synthesis kernels (C ABI via waveformer_b200.ops) vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import haar, haar_c

from helpers import seeded_randn

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-6   # |err| relative to max |value|: one butterfly in fp32
BF16_TOL = 8e-3   # bf16 storage: 2^-8 relative rounding of inputs and outputs


def _ops():
    from waveformer_b200 import ops
    return ops


def _oracle_stack(x):
    c = haar.wavedec3(x.double(), level=1)
    return haar.details_to_stack(c[0], c[1])  # [..., 8, d, h, w]


@pytest.mark.parametrize("shape", [(2, 3, 16, 16, 16), (1, 5, 8, 4, 24), (1, 2, 6, 10, 2), (3, 4, 2, 2, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwt_ncdhw_matches_oracle(shape, dtype):
    ops = _ops()
    x = seeded_randn(shape, 11).to(dtype)
    ll, hf = ops.dwt3d(x.cuda())
    want = _oracle_stack(x.float())
    got = torch.cat([ll.unsqueeze(0), hf], 0).float().cpu().movedim(0, -4)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert float((got.double() - want).abs().max() / want.abs().max()) < tol
    assert hf.shape == (7,) + tuple(ll.shape)


@pytest.mark.parametrize("shape", [(2, 16, 16, 16, 48), (1, 4, 8, 6, 96), (1, 2, 2, 2, 8), (2, 4, 4, 4, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwt_ndhwc_matches_oracle(shape, dtype):
    ops = _ops()
    x = seeded_randn(shape, 12).to(dtype)
    ll, hf = ops.dwt3d_channels_last(x.cuda())
    want = _oracle_stack(x.float().permute(0, 4, 1, 2, 3))          # [B, C, 8, d, h, w]
    got = torch.cat([ll.unsqueeze(0), hf], 0).float().cpu()          # [8, B, d, h, w, C]
    got = got.permute(1, 5, 0, 2, 3, 4)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    assert float((got.double() - want).abs().max() / want.abs().max()) < tol
    ll2, none = ops.dwt3d_channels_last(x.cuda(), need_hf=False)
    assert none is None and torch.equal(ll2, ll)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.bfloat16, 2e-2)])
def test_roundtrip_both_layouts(dtype, tol):
    ops = _ops()
    x = seeded_randn((2, 6, 16, 8, 32), 13).to(dtype).cuda()
    ll, hf = ops.dwt3d(x)
    rec = ops.idwt3d(ll, hf)
    assert float((rec.float() - x.float()).abs().max() / x.float().abs().max()) < tol
    xc = seeded_randn((2, 16, 8, 32, 24), 14).to(dtype).cuda()
    ll, hf = ops.dwt3d_channels_last(xc)
    rec = ops.idwt3d_channels_last(ll, hf)
    assert float((rec.float() - xc.float()).abs().max() / xc.float().abs().max()) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reconstruction_indexing_is_bit_exact(dtype):
    """Integer ramp (exactly representable): DWT -> IDWT must return every voxel to its own index (north star:
    'bit-exact DWT->IDWT reconstruction indexing').  Values are small integers so bf16 holds them exactly too."""
    ops = _ops()
    idx = (torch.arange(2 * 3 * 8 * 8 * 16) % 64).reshape(2, 3, 8, 8, 16).float()
    x = idx.to(dtype).cuda()
    rec = ops.idwt3d(*ops.dwt3d(x))
    assert torch.equal(rec.float().round().cpu(), idx)
    xc = idx.permute(0, 2, 3, 4, 1).contiguous().to(dtype).cuda()
    rec = ops.idwt3d_channels_last(*ops.dwt3d_channels_last(xc))
    assert torch.equal(rec.float().round().cpu(), idx.permute(0, 2, 3, 4, 1))


def test_idwt_matches_oracle_and_c_port():
    ops = _ops()
    c = seeded_randn((2, 3, 8, 4, 6, 8), 15)
    want = haar_c.idwt3d(c.numpy())
    got = ops.idwt3d(c[:, :, 0].contiguous().cuda(), c[:, :, 1:].permute(2, 0, 1, 3, 4, 5).contiguous().cuda())
    assert np.abs(got.cpu().numpy() - want).max() < 5e-6


def test_idwt_gate_and_zero_details():
    ops = _ops()
    ll = seeded_randn((1, 4, 4, 4, 16), 16).cuda()
    hf = seeded_randn((7, 1, 4, 4, 4, 16), 17).cuda()
    gate = torch.sigmoid(seeded_randn((7, 1, 4, 4, 4, 16), 18)).cuda()
    a = ops.idwt3d_channels_last(ll, hf, gate)
    b = ops.idwt3d_channels_last(ll, hf * gate)
    assert float((a - b).abs().max()) < 1e-6
    z = ops.idwt3d_channels_last(ll, None)
    zz = ops.idwt3d_channels_last(ll, torch.zeros_like(hf))
    assert torch.equal(z, zz)


def test_idwt_writes_into_concat_buffer():
    ops = _ops()
    ll = seeded_randn((2, 4, 4, 4, 24), 19).bfloat16().cuda()
    hf = seeded_randn((7, 2, 4, 4, 4, 24), 20).bfloat16().cuda()
    dense = ops.idwt3d_channels_last(ll, hf)
    buf = torch.full((2, 8, 8, 8, 48), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.idwt3d_channels_last(ll, hf, None, buf[..., :24])
    assert torch.equal(buf[..., :24], dense) and bool((buf[..., 24:] == 7).all())


def test_multilevel_helper_api_matches_oracle():
    from waveformer_b200.network_models import WaveletTransform3D
    from waveformer_b200.network_models.wave_helper import waverec3
    x = seeded_randn((2, 4, 32, 16, 16), 21)
    yl, yh = WaveletTransform3D()(x.cuda(), 3)
    want = haar.wavedec3(x, level=3)
    assert len(yh) == 3 and list(yh[0].keys()) == list(haar.DETAIL_KEYS)
    assert float((yl.cpu() - want[0]).abs().max()) < 1e-5
    for lvl in range(3):
        for k in haar.DETAIL_KEYS:
            assert yh[lvl][k].shape == want[1 + lvl][k].shape
            assert float((yh[lvl][k].cpu() - want[1 + lvl][k]).abs().max()) < 1e-5
    assert float((waverec3((yl,) + yh).cpu() - x).abs().max()) < 1e-5


def test_error_conventions():
    ops = _ops()
    with pytest.raises(ValueError):
        ops.dwt3d(torch.zeros(1, 3, 4, 4, device="cuda"))            # odd extent
    with pytest.raises(ValueError):
        ops.dwt3d(torch.zeros(1, 4, 4, 4, device="cuda", dtype=torch.float64))       # unsupported dtype (ptwt: ValueError too)
    with pytest.raises(RuntimeError):
        ops.dwt3d(torch.zeros(1, 4, 4, 4))                            # CPU tensor: no fallback
    from waveformer_b200.network_models import WaveletTransform3D
    with pytest.raises(ValueError):
        WaveletTransform3D(wavelet="db2")


def test_autograd_is_the_adjoint():
    ops = _ops()
    x = seeded_randn((1, 2, 8, 8, 8), 22).cuda().requires_grad_(True)
    ll, hf = ops.dwt3d(x)
    g_ll, g_hf = torch.randn_like(ll), torch.randn_like(hf)
    (ll * g_ll).sum().add((hf * g_hf).sum()).backward()
    want = ops.idwt3d(g_ll, g_hf)
    assert float((x.grad - want).abs().max()) < 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_full_size_roundtrip_properties(dtype):
    """BASELINE config 2 size (2x48x128^3): size-independent properties instead of an oracle run - Parseval and
    reconstruction, plus a sampled comparison against the closed form."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 48, 128, 128, 128), device="cuda", generator=g, dtype=torch.float32).to(dtype)
    ll, hf = ops.dwt3d(x)
    e_in = float(x.float().pow(2).sum(dtype=torch.float64))
    e_out = float(ll.float().pow(2).sum(dtype=torch.float64)) + float(hf.float().pow(2).sum(dtype=torch.float64))
    assert abs(e_in - e_out) < (1e-5 if dtype == torch.float32 else 2e-3) * e_in
    rec = ops.idwt3d(ll, hf)
    err = float((rec.float() - x.float()).abs().max()) / float(x.float().abs().max())
    # north star: 16-bit reconstruction within 2e-2 relative (two roundings of O(|x|) coefficients); fp32 at round-off level
    assert err < {torch.float32: 2e-6, torch.bfloat16: 2e-2, torch.float16: 2.5e-3}[dtype], err
    sub = x[1, 7, 32:40, 64:72, 96:112].float().cpu()
    want = haar.haar_cell_forward(sub.double())
    got = torch.cat([ll[1, 7, 16:20, 32:36, 48:56].unsqueeze(0), hf[:, 1, 7, 16:20, 32:36, 48:56]], 0).float().cpu()
    assert float((got.double() - want).abs().max()) < {torch.float32: 1e-5, torch.bfloat16: 5e-2, torch.float16: 8e-3}[dtype]
