"""Random draws from an integer key (Model.acquire_key).  The reference threads
threefry keys through jax.random; those streams are not reproducible without JAX
(documented differences), so draws are distribution-equivalent only."""
from __future__ import annotations

import numpy as np
import torch

from lib._device import device


def _gen(key) -> torch.Generator:
    g = torch.Generator(device=device())
    g.manual_seed(int(np.asarray(key).reshape(-1)[-1]) & 0x7FFFFFFFFFFFFFFF)
    return g


def normal(key, shape) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=_gen(key), dtype=torch.float64, device=device())


def uniform(key, shape, minval: float = 0.0, maxval: float = 1.0) -> torch.Tensor:
    u = torch.rand(tuple(shape), generator=_gen(key), dtype=torch.float64, device=device())
    return minval + (maxval - minval) * u


def randint(key, minval: int, maxval: int) -> int:
    return int(torch.randint(int(minval), int(maxval), (1,), generator=_gen(key), device=device()).item())
