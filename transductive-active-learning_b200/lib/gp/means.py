"""Prior mean functions (reference: lib/gp/means.py:8-29)."""
from abc import ABC, abstractmethod

import torch

from lib._device import as_f64


class Mean(ABC):
    @abstractmethod
    def __call__(self, x):
        pass

    def vector(self, X) -> torch.Tensor:
        X = as_f64(X)
        return torch.stack([as_f64(self(x)).reshape(()) for x in X])


class ZeroMean(Mean):
    def __call__(self, x):
        return 0

    def vector(self, X) -> torch.Tensor:
        X = as_f64(X)
        return torch.zeros((X.shape[0],), dtype=torch.float64, device=X.device)


class ConstantMean(Mean):
    def __init__(self, value):
        self.value = value

    def __call__(self, x):
        return self.value

    def vector(self, X) -> torch.Tensor:
        X = as_f64(X)
        return torch.full((X.shape[0],), float(self.value), dtype=torch.float64, device=X.device)
