"""Black box given by a Python callable (reference: lib/function/unknown.py:8-28)."""
import torch

from lib._device import as_f64
from lib.function import Function


class UnknownFunction(Function):
    def __init__(self, q: int, f, use_objective_as_constraint: bool = False):
        self.q = q
        self.first_constraint = 0 if use_objective_as_constraint else 1
        self._f = f

    def observe(self, X) -> torch.Tensor:
        """[q, m]: f evaluated point by point (vmap over the rows of X, :26-28)."""
        X = as_f64(X)
        return torch.stack([as_f64(self._f(X[j])).reshape(-1) for j in range(X.shape[0])]).T.contiguous()
