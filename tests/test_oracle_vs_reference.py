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


def test_mirror_tta_restatement_matches_reference_predictor():
    """oracle/prediction.py vs the live ``light_training.prediction.Predictor`` (its heavy optional imports stubbed)."""
    import importlib
    import sys
    import types
    for name in ("SimpleITK", "skimage", "skimage.measure", "light_training.preprocessing.resampling.default_resampling"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["light_training.preprocessing.resampling.default_resampling"].resample_data_or_seg_to_shape = lambda *a, **k: None
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    pred = importlib.import_module("light_training.prediction")
    from oracle import prediction as op
    x = torch.randn(1, 2, 12, 10, 8, generator=torch.Generator().manual_seed(3))

    def net(v):
        return v * torch.arange(8.0)[None, None, None, None, :] + 0.1 * v.flip(2)

    class _Id(torch.nn.Module):
        def forward(self, v):
            return v

    for axes in ([0, 1, 2], [1], None):
        ref = pred.Predictor(lambda v, model, **k: net(v), mirror_axes=axes).maybe_mirror_and_predict(x, _Id(), torch.device("cpu"))
        assert float((ref - op.mirror_and_predict(x, net, axes)).abs().max()) == 0.0
    props = {"shape_after_cropping_before_resample": (15, 9, 11)}
    want = pred.Predictor.predict_raw_probability(x.clone(), props)
    assert float((want.float() - op.predict_raw_probability(x, (15, 9, 11)).float()).abs().max()) == 0.0
