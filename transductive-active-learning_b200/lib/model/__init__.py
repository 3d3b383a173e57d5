"""Statistical model base class (reference: lib/model/__init__.py:10-69)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np
import torch

from lib.noise import Noise


def _split_key(key):
    """Deterministic stand-in for jax.random.split on an integer key
    (splitmix64).  The reference threads threefry keys; that PRNG is not
    reproducible without JAX (DESIGN.md "Randomness")."""
    z = (int(key) + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    a = z
    a = ((a ^ (a >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    a = ((a ^ (a >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    a = a ^ (a >> 31)
    return z, a


class Model(ABC):
    q: int
    first_constraint: int
    noise_rate: Noise

    def __init__(self, key, q: int, beta, noise_rate: Noise,
                 use_objective_as_constraint: bool = False):
        self._key = 0 if key is None else int(np.asarray(key).reshape(-1)[-1])
        self.q = q
        self.first_constraint = 0 if use_objective_as_constraint else 1
        self.beta = (lambda t: beta) if not callable(beta) else beta
        self.noise_rate = noise_rate

    def acquire_key(self):
        self._key, key = _split_key(self._key)
        return key

    @property
    def constraint_set(self) -> torch.Tensor:
        return torch.arange(self.q) >= self.first_constraint

    @property
    @abstractmethod
    def t(self) -> int:
        pass

    @abstractmethod
    def step(self, X, y):
        pass
