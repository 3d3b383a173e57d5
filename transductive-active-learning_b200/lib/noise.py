"""Observation-noise models (reference: lib/noise.py:19-46).  `at(X)` returns
the [q, n] noise standard deviations that every update / scoring kernel takes
as a per-point array (heteroscedastic noise included)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, List

import torch

from lib._device import as_f64


class Noise(ABC):
    @abstractmethod
    def at(self, X) -> torch.Tensor:
        """Noise rate at every point in `X`: [q, n]."""


class HomoscedasticNoise(Noise):
    """lib/noise.py:28-35."""

    def __init__(self, q: int, noise_rates):
        self.q = q
        self._noise_rates = as_f64(noise_rates).reshape(-1)

    def at(self, X) -> torch.Tensor:
        n = X.shape[0]
        return self._noise_rates[:, None].expand(self.q, n).contiguous()


class HeteroscedasticNoise(Noise):
    """lib/noise.py:38-46."""

    def __init__(self, noise_rates: List[Callable]):
        self.q = len(noise_rates)
        self._noise_rates = noise_rates

    def at(self, X) -> torch.Tensor:
        return torch.stack([as_f64(self._noise_rates[i](X)).reshape(-1) for i in range(self.q)])
