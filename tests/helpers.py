"""Shared helpers for the tests (seeded inputs identical to scripts/make_golden.py)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seeded_randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def checksum(t: torch.Tensor) -> float:
    return float(t.double().sum())


def sample_positions(numel: int, count: int, seed: int) -> np.ndarray:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (count,), generator=g).numpy().astype(np.int64)


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name))


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def sub_state(sd, prefix):
    return {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}


def max_rel(a, b) -> float:
    """max |a-b| / max |b|  (the north star's 'max-relative' error)."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_input_matches(x: torch.Tensor, expected_sum: float):
    got = checksum(x)
    assert abs(got - float(expected_sum)) <= 1e-6 * max(1.0, abs(float(expected_sum))), (
        "seeded input differs from the one the golden fixture was generated with (torch RNG drift?)")
