"""Oracle self-checks for the Haar restatement (ptwt boundary: parity unpinned -> first-principles KATs)."""
import math

import numpy as np
import pytest
import torch

from oracle import haar, haar_c

S = 1.0 / (2.0 * math.sqrt(2.0))


def test_constant_volume_only_ll():
    x = torch.full((1, 2, 8, 8, 8), 3.0, dtype=torch.float64)
    ll, det = haar.wavedec3(x, level=1)
    assert torch.allclose(ll, torch.full_like(ll, 3.0 * 2.0 * math.sqrt(2.0)))
    for k, v in det.items():
        assert float(v.abs().max()) < 1e-12, k


def test_impulse_gives_hadamard_table():
    had = haar.hadamard8()
    for m in range(8):
        i, j, k = (m >> 2) & 1, (m >> 1) & 1, m & 1
        x = torch.zeros(1, 2, 2, 2, dtype=torch.float64)
        x[0, i, j, k] = 1.0
        c = haar.wavedec3(x, level=1)
        st = haar.details_to_stack(c[0], c[1]).reshape(8)
        assert torch.allclose(st, had[:, m] * S, atol=1e-15)


def test_detail_key_order_and_shapes():
    x = torch.randn(2, 3, 8, 12, 4)
    c = haar.wavedec3(x, level=2)
    assert len(c) == 3 and c[0].shape == (2, 3, 2, 3, 1)
    assert list(c[1].keys()) == list(haar.DETAIL_KEYS) == ["aad", "ada", "add", "daa", "dad", "dda", "ddd"]
    assert c[1]["ddd"].shape == (2, 3, 2, 3, 1) and c[2]["aad"].shape == (2, 3, 4, 6, 2)  # coarsest first


def test_subband_naming_follows_axes():
    # variation along W only -> only 'aad' (last letter = W) is non-zero besides LL
    x = torch.zeros(1, 4, 4, 4, dtype=torch.float64)
    x[..., 0::2] = 1.0
    ll, det = haar.wavedec3(x, level=1)
    nz = [k for k, v in det.items() if float(v.abs().max()) > 1e-12]
    assert nz == ["aad"] and float(det["aad"].min()) > 0  # x[even]-x[odd] > 0 -> hi = (s, -s)
    x = torch.zeros(1, 4, 4, 4, dtype=torch.float64)
    x[:, 0::2] = 1.0
    _, det = haar.wavedec3(x, level=1)
    assert [k for k, v in det.items() if float(v.abs().max()) > 1e-12] == ["daa"]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-14)])
def test_roundtrip_and_parseval(dtype, tol):
    x = torch.randn(2, 5, 16, 8, 12, generator=torch.Generator().manual_seed(0)).to(dtype)
    c = haar.wavedec3(x, level=3 if dtype == torch.float64 else 2)
    rec = haar.waverec3(c)
    assert float((rec - x).abs().max()) < tol * 10
    energy = float(c[0].double().pow(2).sum()) + sum(float(v.double().pow(2).sum()) for d in c[1:] for v in d.values())
    assert abs(energy - float(x.double().pow(2).sum())) < 1e-4 * energy


def test_closed_form_and_c_port_match_conv_form():
    x = torch.randn(3, 2, 6, 8, 10, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    c = haar.wavedec3(x, level=1)
    st = haar.details_to_stack(c[0], c[1])
    assert float((haar.haar_cell_forward(x) - st).abs().max()) < 1e-14
    assert float((haar.haar_cell_inverse(st) - x).abs().max()) < 1e-14
    cc = haar_c.dwt3d(x.numpy(), threads=3)
    assert np.abs(cc - st.numpy()).max() < 1e-14
    assert np.abs(haar_c.idwt3d(cc, threads=2) - x.numpy()).max() < 1e-14
    x32 = x.float()
    assert np.abs(haar_c.dwt3d(x32.numpy()) - haar.haar_cell_forward(x32).numpy()).max() < 1e-5


def test_index_exactness_on_integer_ramp():
    # values exactly representable: reconstruction must be exact up to the 1/(2*sqrt2) scaling round-off
    x = torch.arange(8 * 8 * 8, dtype=torch.float64).reshape(1, 8, 8, 8)
    rec = haar.waverec3(haar.wavedec3(x, level=1))
    assert torch.equal(rec.round(), x)


def test_errors_match_ptwt_conventions():
    with pytest.raises(ValueError):
        haar.wavedec3(torch.zeros(1, 4, 4, 4, dtype=torch.bfloat16))
    with pytest.raises(ValueError):
        haar.wavedec3(torch.zeros(1, 4, 4, 4), wavelet="db2")
    ll, det = haar.wavedec3(torch.zeros(1, 4, 4, 4))
    bad = dict(det)
    bad.pop("ddd")
    with pytest.raises(ValueError):
        haar.waverec3((ll, bad))


def test_odd_extent_zero_padded():
    x = torch.randn(1, 5, 4, 4, dtype=torch.float64)
    ll, det = haar.wavedec3(x, level=1)
    assert ll.shape == (1, 3, 2, 2)
    rec = haar.waverec3((ll, det))
    assert float((rec[:, :5] - x).abs().max()) < 1e-14
