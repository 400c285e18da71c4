"""GPU: the tcgen05 / TMEM attention path (bf16) against the fp32 oracle, the reference golden outputs and the
CUDA-core kernels run on the same inputs (WF_ATTN_IMPL=simt)."""
import os

import pytest
import torch

from oracle import model as om
from oracle.state import ModelConfig, make_state_dict

from helpers import load_npz, max_rel, seeded_randn, sub_state

pytestmark = pytest.mark.gpu
CFG = ModelConfig(img_size=(128,) * 3)


@pytest.fixture(scope="module")
def sd():
    return make_state_dict(CFG, seed=0)


def _attn(sd, stage, c, h):
    from waveformer_b200.network_models import Attention
    m = Attention(c, num_heads=h, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
    m.load_state_dict(sub_state(sd, f"waveformer_encoder.block{stage + 1}.1.attn"), strict=True)
    return m.cuda().to(torch.bfloat16)


def _run(m, x, impl):
    old = os.environ.get("WF_ATTN_IMPL")
    os.environ["WF_ATTN_IMPL"] = impl
    try:
        with torch.no_grad():
            y = m.forward_grid(x)
        torch.cuda.synchronize()
        return y
    finally:
        if old is None:
            os.environ.pop("WF_ATTN_IMPL", None)
        else:
            os.environ["WF_ATTN_IMPL"] = old


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_tensor_core_path_matches_reference_golden(sd, stage, c, h, b_):
    g = load_npz("attention_ws8.npz")
    x = seeded_randn((b_, 512, c), 100 + stage).cuda().bfloat16().reshape(b_, 8, 8, 8, c)
    m = _attn(sd, stage, c, h)
    tc = _run(m, x, "tc").float().cpu().reshape(b_, 512, c)
    simt = _run(m, x, "simt").float().cpu().reshape(b_, 512, c)
    assert torch.isfinite(tc).all()
    assert max_rel(tc, g[f"out_{c}"]) < 2e-2          # bf16 gate vs the fp32 reference
    assert max_rel(tc, simt) < 1.5e-2                  # two independent device implementations agree


@pytest.mark.parametrize("grid,batch", [((16, 16, 16), 2), ((32, 32, 32), 2), ((8, 16, 24), 1), ((64, 64, 64), 1)])
def test_tensor_core_path_many_windows(sd, grid, batch):
    """Several windows per CTA (persistent loop, double-buffered bulk copies) and the reshape-only reverse."""
    p = "waveformer_encoder.block1.0.attn"
    m = _attn(sd, 0, 48, 3)
    m.load_state_dict(sub_state(sd, p), strict=True)
    m = m.cuda().to(torch.bfloat16)
    x = seeded_randn((batch,) + grid + (48,), 35)
    got = _run(m, x.cuda().bfloat16(), "tc").float().cpu()
    if grid[0] * grid[1] * grid[2] <= 16 ** 3 * 2:
        want = om.window_attention(sd, p, om.window_partition(x.bfloat16().float(), 8), 3).reshape(x.shape)
        assert max_rel(got, want) < 2e-2
    simt = _run(m, x.cuda().bfloat16(), "simt").float().cpu()
    assert max_rel(got, simt) < 1.5e-2


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_fp16_operands_fp32_stream(sd, stage, c, h, b_):
    """The precision policy of prepare_inference: fp32 activations in, fp16 tensor-core operands (10-bit mantissa),
    fp32 result.  Must be ~8x closer to the fp32 reference than the bf16 gate (tolerance 3e-3 vs 2e-2)."""
    from waveformer_b200.network_models import Attention
    g = load_npz("attention_ws8.npz")
    m = Attention(c, num_heads=h, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
    m.load_state_dict(sub_state(sd, f"waveformer_encoder.block{stage + 1}.1.attn"), strict=True)
    m = m.cuda()                                    # fp32 master weights; the fp16 copies are cached by ops.cast_cached
    m.compute_dtype, m.out_dtype = torch.float16, torch.float32
    x = seeded_randn((b_, 512, c), 100 + stage).cuda().reshape(b_, 8, 8, 8, c)
    y = _run(m, x, "tc")
    assert y.dtype == torch.float32
    assert max_rel(y.cpu().reshape(b_, 512, c), g[f"out_{c}"]) < 3e-3
    m.compute_dtype, m.out_dtype = torch.bfloat16, torch.float32
    y16 = _run(m, x, "tc")
    assert y16.dtype == torch.float32 and max_rel(y16.cpu().reshape(b_, 512, c), g[f"out_{c}"]) < 2e-2


@pytest.mark.parametrize("stage,c,h,b_", [(0, 48, 3, 3), (1, 96, 6, 2), (2, 192, 12, 1), (3, 384, 24, 1)])
def test_compensated_fp16_operands(sd, stage, c, h, b_):
    """attention="fp16x2" of the precision policy: hi / lo fp16 pairs for x, the weights, q, k and O (three tcgen05.mma per
    product).  The scores are then exact to fp32 level and what is left is the fp16 rounding of P and v: an order of
    magnitude closer to the fp32 reference than plain fp16 operands (tolerance 4e-4 vs 3e-3)."""
    from waveformer_b200.network_models import Attention
    g = load_npz("attention_ws8.npz")
    m = Attention(c, num_heads=h, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
    m.load_state_dict(sub_state(sd, f"waveformer_encoder.block{stage + 1}.1.attn"), strict=True)
    m = m.cuda()
    m.compute_dtype, m.out_dtype, m.split_operands = torch.float16, torch.float32, True
    x = seeded_randn((b_, 512, c), 100 + stage).cuda().reshape(b_, 8, 8, 8, c)
    y = _run(m, x, "tc")
    assert y.dtype == torch.float32 and torch.isfinite(y).all()
    err = max_rel(y.cpu().reshape(b_, 512, c), g[f"out_{c}"])
    m.split_operands = False
    plain = max_rel(_run(m, x, "tc").cpu().reshape(b_, 512, c), g[f"out_{c}"])
    assert err < 4e-4 and err < 0.5 * plain, (err, plain)


@pytest.mark.parametrize("grid,batch", [((32, 32, 32), 2), ((8, 16, 24), 1)])
def test_compensated_path_many_windows(sd, grid, batch):
    """Persistent loop of the compensated core: single Q / K buffer refilled after the score MMAs, V double-buffered."""
    from waveformer_b200.network_models import Attention
    p = "waveformer_encoder.block1.0.attn"
    m = Attention(48, num_heads=3, qkv_bias=True, window_size=8, img_size=(8, 8, 8)).eval()
    m.load_state_dict(sub_state(sd, p), strict=True)
    m = m.cuda()
    m.compute_dtype, m.out_dtype, m.split_operands = torch.float16, torch.float32, True
    x = seeded_randn((batch,) + grid + (48,), 36)
    got = _run(m, x.cuda(), "tc").cpu()
    simt = _run(m, x.cuda(), "simt").cpu()             # fp32 CUDA-core kernels on the same input
    assert max_rel(got, simt) < 4e-4
