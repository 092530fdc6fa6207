"""Shared test helpers: golden loading, oracle parameter construction, comparison metrics."""
from __future__ import annotations

import os

import numpy as np
import torch

from cases import Case, make_state_dict
from oracle import snn_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    return dict(np.load(os.path.join(GOLD, f"actor_{name}.npz")))


def oracle_params(c: Case) -> O.ActorParams:
    return O.ActorParams.from_state_dict(make_state_dict(c), spike_ts=c.spike_ts,
                                         enc_eps=c.enc_eps, dec_tanh=c.dec_tanh)


def unpack_bits(bits: np.ndarray, shape) -> torch.Tensor:
    n = int(np.prod(shape))
    return torch.from_numpy(np.unpackbits(bits)[:n].reshape(shape).astype(np.float32))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|_inf, tiny): relative to the tensor's scale (the 1e-4 bar of BASELINE.json)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def mismatch_rate(a: torch.Tensor, b: torch.Tensor) -> float:
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float((a.to(torch.float32) != b.to(torch.float32)).to(torch.float64).mean())
