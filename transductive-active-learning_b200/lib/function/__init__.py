"""Wrapper around the unknown function (reference: lib/function/__init__.py:5-16)."""
from abc import ABC, abstractmethod


class Function(ABC):
    q: int
    first_constraint: int

    @abstractmethod
    def observe(self, X):
        """Noisy observation of f at X: [q, m]."""


class TableFunction(Function):
    """Black box backed by a table of pre-drawn noisy observations
    y_table[q, N] over the finite domain, resident in HBM.  `observe` is a
    lookup, so a benchmark / parity harness sees identical observations on
    every implementation and the BO loop needs no host round trip."""

    device_resident = True  # `observe` never leaves the device: the greedy step stays one fused call

    def __init__(self, model_domain, y_table, first_constraint: int = 1):
        from lib._device import as_f64
        self.domain = as_f64(model_domain)
        self.y_table = as_f64(y_table)
        self.q = self.y_table.shape[0]
        self.first_constraint = first_constraint

    def observe(self, X):
        from lib.utils import get_indices
        idx = get_indices(self.domain, X)
        return self.y_table[:, idx]
