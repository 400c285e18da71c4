"""Pin the oracle (oracle/) against fixtures produced by the UNMODIFIED reference (scripts/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import sliding_window as osw
from oracle.state import ModelConfig, make_state_dict, relative_position_index, spec_as_json, state_spec

from helpers import (assert_input_matches, load_json, load_npz, max_rel, sample_positions, seeded_randn, sub_state)

CFG = ModelConfig(img_size=(128,) * 3)


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(CFG, seed=0)


@pytest.mark.parametrize("img", [64, 128])
def test_state_spec_matches_reference_dump(img):
    assert spec_as_json(state_spec(ModelConfig(img_size=(img,) * 3))) == load_json(f"state_dict_spec_{img}.json")


def test_relative_position_index_matches_reference():
    ref = load_npz("relative_position_index_ws8.npz")["index"].astype(np.int64)
    idx = relative_position_index(8).numpy()
    assert np.array_equal(idx, ref)
    assert idx.min() == 0 and idx.max() == 546 and len(np.unique(idx)) == 547  # SURVEY 8c known answer


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_attention(sd, stage, c, h, b_):
    g = load_npz("attention_ws8.npz")
    x = seeded_randn((b_, 512, c), 100 + stage)
    assert_input_matches(x, g[f"in_sum_{c}"])
    y = om.window_attention(sd, f"waveformer_encoder.block{stage + 1}.1.attn", x, h)
    assert max_rel(y, g[f"out_{c}"]) < 2e-5


def test_attention_zero_qkv_is_mean_of_v(sd):
    # known answer: zero qkv weight and zero table -> uniform softmax -> proj(mean(v)) = proj(bias_v)
    p = "waveformer_encoder.block1.0.attn"
    local = dict(sd)
    local[f"{p}.qkv.weight"] = torch.zeros_like(sd[f"{p}.qkv.weight"])
    local[f"{p}.relative_position_bias_table"] = torch.zeros_like(sd[f"{p}.relative_position_bias_table"])
    x = seeded_randn((1, 512, 48), 7)
    y = om.window_attention(local, p, x, 3)
    v = sd[f"{p}.qkv.bias"][96:]
    want = torch.nn.functional.linear(v, sd[f"{p}.proj.weight"], sd[f"{p}.proj.bias"])
    assert float((y - want).abs().max()) < 1e-5


def test_block_stage1(sd):
    g = load_npz("block_stage1.npz")
    x = seeded_randn((1, 64, 64, 64, 48), 200)
    assert_input_matches(x, g["in_sum"])
    y, hf = om.block(sd, "waveformer_encoder.block1.1", x, heads=3, level=3, ws=8)
    assert max_rel(y.reshape(-1)[g["pos"]], g["out"]) < 2e-5
    assert abs(float(y.double().sum()) - float(g["out_sum"])) < 1e-4 * float(g["out_abs_sum"])
    assert len(hf) == 3
    for li, d in enumerate(hf):
        for key, t in d.items():
            assert tuple(t.shape) == tuple(g[f"hf{li}_{key}_shape"])
            p = sample_positions(t.numel(), 512, 300 + li)
            assert max_rel(t.reshape(-1)[p], g[f"hf{li}_{key}"]) < 2e-5


def test_patch_merging(sd):
    g = load_npz("patch_merging.npz")
    x = seeded_randn((1, 8, 8, 8, 48), 400)
    assert_input_matches(x, g["in_sum"])
    assert max_rel(om.patch_merging(sd, "waveformer_encoder.downsample_1", x), g["out"]) < 2e-5


def test_idwt_block(sd):
    g = load_npz("idwt_block.npz")
    inp = seeded_randn((1, 384, 4, 4, 4), 500)
    skip = seeded_randn((1, 96, 16, 16, 16), 501)
    assert_input_matches(torch.cat([inp.reshape(-1), skip.reshape(-1)]), g["in_sum"])
    keys = ("aad", "ada", "add", "daa", "dad", "dda", "ddd")
    hf = ({k: seeded_randn((1, 96, 4, 4, 4), 510 + i) for i, k in enumerate(keys)},
          {k: seeded_randn((1, 96, 8, 8, 8), 520 + i) for i, k in enumerate(keys)})
    assert max_rel(om.idwt_block(sd, "decoder3", inp, skip, hf), g["out"]) < 2e-5


def test_waveformer_forward_128(sd):
    """BASELINE config 1 (1x4x128^3 fp32 on CPU) against the reference's logits."""
    g = load_npz("waveformer_128.npz")
    x = seeded_randn((1, 4, 128, 128, 128), 1)
    assert_input_matches(x, g["in_sum"])
    with torch.no_grad():
        y, mid = om.waveformer_forward(sd, x, CFG, return_intermediates=True)
    assert max_rel(y.reshape(-1)[g["pos"]], g["logits"]) < 1e-4
    assert np.array_equal(np.bincount(y.argmax(1).reshape(-1).numpy(), minlength=4), g["label_hist"]) or \
        np.abs(np.bincount(y.argmax(1).reshape(-1).numpy(), minlength=4) - g["label_hist"]).sum() < 50
    e = load_npz("encoder_128.npz")
    for i, o in enumerate(mid["outs"]):
        assert max_rel(o.reshape(-1)[sample_positions(o.numel(), 2048, 10 + i)], e[f"out{i}"]) < 5e-5
    for si, hfs in enumerate(mid["outs_hf"]):
        for li, d in enumerate(hfs):
            t = d["dad"]
            assert max_rel(t.reshape(-1)[sample_positions(t.numel(), 512, 20 + 4 * si + li)], e[f"hf_s{si}_l{li}_dad"]) < 3e-4  # details are differences of O(1) values: fp32 cancellation


def test_window_geometry_240x240x155():
    g = load_json("windows_240x240x155.json")
    iv = osw.scan_interval((240, 240, 155), (128, 128, 128), 0.5)
    assert list(iv) == g["interval"] == [64, 64, 64]
    starts = osw.window_starts((240, 240, 155), (128, 128, 128), iv)
    assert [list(s) for s in starts] == g["starts"] and len(starts) == 18
    m = osw.importance_map((128, 128, 128), "gaussian", 0.125)
    assert abs(float(m.min()) - g["imap_min"]) < 1e-9 and abs(float(m.max()) - g["imap_max"]) < 1e-7
    assert abs(float(m.double().sum()) - g["imap_sum"]) < 1e-6 * g["imap_sum"]
    assert np.allclose([float(m[i, i, i]) for i in range(0, 128, 8)], g["imap_diag"], rtol=1e-6)


@pytest.mark.parametrize("name,shape,roi,ov,mode,bs", [
    ("a", (1, 2, 40, 36, 30), (16, 16, 16), 0.5, "gaussian", 2),
    ("b", (2, 2, 20, 33, 17), (16, 16, 16), 0.25, "gaussian", 3),
    ("c", (1, 2, 12, 40, 16), (16, 16, 16), 0.5, "constant", 4),  # first axis smaller than the roi -> padded
])
def test_sliding_window_small(name, shape, roi, ov, mode, bs):
    g = load_npz("sliding_window_small.npz")
    wconv = seeded_randn((3, 2, 3, 3, 3), 600) * 0.2
    x = seeded_randn(shape, 610 + ord(name))
    assert_input_matches(x, g[f"{name}_in_sum"])
    y = osw.sliding_window_inference(x, roi, bs, lambda p: torch.nn.functional.conv3d(p, wconv, padding=1), ov, mode)
    assert tuple(y.shape) == tuple(g[f"{name}_out"].shape)
    assert max_rel(y, g[f"{name}_out"]) < 1e-5


def test_sliding_window_identity_network_returns_input():
    x = seeded_randn((1, 3, 40, 24, 30), 9)
    y = osw.sliding_window_inference(x, (16, 16, 16), 2, lambda p: p, 0.5, "gaussian")
    assert float((y - x).abs().max()) < 1e-5


@pytest.mark.parametrize("name,kw", [("default", dict(to_onehot_y=True, softmax=True)),
                                     ("nobg_sq", dict(to_onehot_y=True, softmax=True, include_background=False, squared_pred=True,
                                                      lambda_dice=0.7, lambda_ce=1.3)),
                                     ("jaccard_batch", dict(to_onehot_y=True, softmax=True, jaccard=True, batch=True))])
def test_dice_ce_loss_matches_reference_golden(name, kw):
    """waveformer_b200.losses.DiceCELoss (BASELINE configs[4]: Dice + CE) vs the reference's vendored MONAI loss
    (scripts/make_golden_loss.py), value and gradient; also under bf16 logits (autocast training)."""
    from waveformer_b200.losses import DiceCELoss
    g = load_npz("dice_ce_loss.npz")
    x = (seeded_randn((2, 4, 12, 10, 14), 40) * 2.0).requires_grad_(True)
    y = torch.randint(0, 4, (2, 1, 12, 10, 14), generator=torch.Generator().manual_seed(41))
    loss = DiceCELoss(**kw)(x, y)
    loss.backward()
    assert abs(float(loss) - float(g[f"{name}_loss"])) < 1e-6
    assert float((x.grad - torch.from_numpy(g[f"{name}_grad"])).abs().max()) < 1e-7
    lb = DiceCELoss(**kw)(x.detach().bfloat16(), y)
    assert lb.dtype == torch.float32 and abs(float(lb) - float(g[f"{name}_loss"])) < 2e-2
