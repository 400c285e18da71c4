"""Authoring-container check: oracle vs the UNMODIFIED reference executed from /root/reference (skipped on the GPU box)."""
import pytest
import torch

from oracle import model as om
from oracle import ref_harness as rh
from oracle import sliding_window as osw
from oracle.state import ModelConfig, make_state_dict, state_spec

from helpers import max_rel, seeded_randn

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    return rh.load_reference()


def test_forward_64_matches_reference(ref):
    cfg = ModelConfig(img_size=(64,) * 3)
    m = ref.Waveformer(**cfg.kwargs()).eval()
    assert [k for k, _, _ in state_spec(cfg)] == list(m.state_dict().keys())
    sd = make_state_dict(cfg, seed=3)
    m.load_state_dict(sd, strict=True)
    x = seeded_randn((2, 4, 64, 64, 64), 5)
    with torch.no_grad():
        assert max_rel(om.waveformer_forward(sd, x, cfg), m(x)) < 5e-5


def test_wavelet_helper_call_site(ref):
    # WaveletTransform3D.forward (wave_helper.py:349-353): (Yl, tuple of dicts, coarsest first)
    w = ref.WaveletTransform3D(wavelet="db1", mode="zero")
    x = seeded_randn((1, 2, 16, 16, 16), 6)
    yl, yh = w(x, 2)
    assert yl.shape == (1, 2, 4, 4, 4) and len(yh) == 2 and yh[0]["aad"].shape == (1, 2, 4, 4, 4)


def test_sliding_window_with_model(ref):
    cfg = ModelConfig(img_size=(64,) * 3)
    sd = make_state_dict(cfg, seed=1)
    m = ref.Waveformer(**cfg.kwargs()).eval()
    m.load_state_dict(sd, strict=True)
    Inferer = rh.load_reference_inferer()
    x = seeded_randn((1, 4, 96, 80, 70), 8)
    with torch.no_grad():
        want = Inferer(roi_size=(64, 64, 64), sw_batch_size=2, overlap=0.5, mode="gaussian")(x, m)
        got = osw.sliding_window_inference(x, (64, 64, 64), 2, lambda p: om.waveformer_forward(sd, p, cfg), 0.5, "gaussian")
    assert max_rel(got, want) < 5e-5
