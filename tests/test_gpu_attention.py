"""GPU parity: window attention (partition + attention + reshape-only reverse) vs the oracle and the golden fixtures."""
import pytest
import torch

from oracle import model as om
from oracle.state import ModelConfig, make_state_dict

from helpers import assert_input_matches, load_npz, max_rel, seeded_randn, sub_state

pytestmark = pytest.mark.gpu
CFG = ModelConfig(img_size=(128,) * 3)


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(CFG, seed=0)


def _module(sd, stage, c, h, dtype):
    from waveformer_b200.network_models import Attention
    m = Attention(c, num_heads=h, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
    m.load_state_dict(sub_state(sd, f"waveformer_encoder.block{stage + 1}.1.attn"), strict=True)
    return m.cuda().to(dtype)


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_attention_fp32_matches_reference_golden(sd, stage, c, h, b_):
    g = load_npz("attention_ws8.npz")
    x = seeded_randn((b_, 512, c), 100 + stage)
    assert_input_matches(x, g[f"in_sum_{c}"])
    with torch.no_grad():
        y = _module(sd, stage, c, h, torch.float32)(x.cuda())
    assert max_rel(y.cpu(), g[f"out_{c}"]) < 2e-5       # fp32 gate of the north star is 1e-4 on logits


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_attention_bf16_matches_reference_golden(sd, stage, c, h, b_):
    g = load_npz("attention_ws8.npz")
    x = seeded_randn((b_, 512, c), 100 + stage)
    with torch.no_grad():
        y = _module(sd, stage, c, h, torch.bfloat16)(x.cuda().bfloat16())
    assert max_rel(y.float().cpu(), g[f"out_{c}"]) < 2e-2  # bf16 gate of the north star


@pytest.mark.parametrize("grid,ws", [((16, 16, 16), 8), ((8, 16, 24), 8), ((8, 8, 8), 4)])
def test_partition_and_reshape_only_reverse(sd, grid, ws):
    """Several windows per volume: the kernel must reproduce window_partition's order on the way in and the
    reference's reshape-only (scrambling) 'reverse' on the way out."""
    p = "waveformer_encoder.block1.0.attn"
    local = dict(sd)
    if ws != 8:
        from oracle.state import relative_position_index
        local[f"{p}.relative_position_index"] = relative_position_index(ws)
        local[f"{p}.relative_position_bias_table"] = seeded_randn(((2 * ws - 1) ** 3, 3), 31) * 0.5
    from waveformer_b200.network_models import Attention
    m = Attention(48, num_heads=3, qkv_bias=True, window_size=ws).eval()
    m.load_state_dict(sub_state(local, p), strict=True)
    m = m.cuda()
    x = seeded_randn((2,) + grid + (48,), 32)
    want = om.window_attention(local, p, om.window_partition(x, ws), 3).reshape(x.shape)
    with torch.no_grad():
        got = m.forward_grid(x.cuda())
    assert max_rel(got.cpu(), want) < 2e-5


def test_bias_cache_follows_table_updates(sd):
    m = _module(sd, 0, 48, 3, torch.float32)
    x = seeded_randn((1, 512, 48), 33).cuda()
    with torch.no_grad():
        a = m(x)
        m.relative_position_bias_table.mul_(0.0)
        b = m(x)
    assert float((a - b).abs().max()) > 1e-4


def test_zero_qkv_weight_known_answer(sd):
    m = _module(sd, 0, 48, 3, torch.float32)
    with torch.no_grad():
        m.qkv.weight.zero_()
        m.relative_position_bias_table.zero_()
        y = m(seeded_randn((1, 512, 48), 34).cuda())
        want = torch.nn.functional.linear(m.qkv.bias[96:], m.proj.weight, m.proj.bias)
    assert float((y - want).abs().max()) < 1e-5


@pytest.mark.parametrize("grid,ws,c,h", [((16, 16, 16), 8, 48, 3), ((8, 16, 8), 8, 96, 6), ((8, 8, 8), 4, 48, 3),
                                        ((8, 8, 8), 8, 64, 2)])
def test_attention_gradients_match_oracle_autograd(sd, grid, ws, c, h):
    """wf_window_attn_bwd (attention core) + the Linear gradients vs fp64 autograd through the oracle restatement,
    including the relative-position table gradient and the reshape-only reverse on the way back."""
    from oracle.state import relative_position_index
    from waveformer_b200.network_models import Attention
    p = "attn"
    hd = c // h
    ref = {f"{p}.qkv.weight": seeded_randn((3 * c, c), 41) * c ** -0.5, f"{p}.qkv.bias": seeded_randn((3 * c,), 42) * 0.1,
           f"{p}.proj.weight": seeded_randn((c, c), 43) * c ** -0.5, f"{p}.proj.bias": seeded_randn((c,), 44) * 0.1,
           f"{p}.relative_position_bias_table": seeded_randn(((2 * ws - 1) ** 3, h), 45) * 0.5,
           f"{p}.relative_position_index": relative_position_index(ws)}
    m = Attention(c, num_heads=h, qkv_bias=True, window_size=ws).train()
    m.load_state_dict(sub_state(ref, p), strict=True)
    m = m.cuda()
    x = seeded_randn((2,) + grid + (c,), 46)
    gout = seeded_randn((2,) + grid + (c,), 47)

    xg = x.cuda().requires_grad_(True)
    y = m.forward_grid(xg)
    y.backward(gout.cuda())
    got = {"x": xg.grad, "qkv.weight": m.qkv.weight.grad, "qkv.bias": m.qkv.bias.grad, "proj.weight": m.proj.weight.grad,
           "proj.bias": m.proj.bias.grad, "relative_position_bias_table": m.relative_position_bias_table.grad}

    ref64 = {k: (v.double().requires_grad_(True) if v.is_floating_point() else v) for k, v in ref.items()}
    x64 = x.double().requires_grad_(True)
    y64 = om.window_attention(ref64, p, om.window_partition(x64, ws), h).reshape(x.shape)
    assert max_rel(y.detach().cpu(), y64.detach().float()) < 2e-5
    y64.backward(gout.double())
    want = {"x": x64.grad}
    want.update({k: ref64[f"{p}.{k}"].grad for k in got if k != "x"})
    for k, g in got.items():
        assert g is not None, k
        assert max_rel(g.cpu(), want[k].float()) < 1e-4, k
    assert hd in (16, 32)
