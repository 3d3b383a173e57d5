"""Kernel base classes (reference: lib/gp/kernels/base.py:12-56).

`cross_covariance(X, Y)` launches the fused distance -> exp Gram kernel
(tal_gram_stationary) and, like the reference's vmap nesting (:17-22), returns
a [len(Y), len(X)] matrix."""
from __future__ import annotations

from abc import ABC

import torch

from lib._device import as_f64


class Kernel(ABC):
    kind: int = -1  # TAL_KERNEL_* of include/tal_b200.h

    def _gram_params(self):
        raise NotImplementedError

    def cross_covariance(self, X, Y, dtype: torch.dtype = torch.float64) -> torch.Tensor:
        import ctypes as C
        from talgp import _abi
        _abi.require_device()
        lib = _abi.load()
        X, Y = as_f64(X), as_f64(Y)
        n, d = X.shape
        m = Y.shape[0]
        ld = (n + 3) // 4 * 4
        out = torch.zeros((m, ld), dtype=dtype, device=X.device)
        variance, lengthscale = self._gram_params()
        _abi.check(lib.tal_gram_stationary(
            self.kind, C.c_void_p(X.data_ptr()), n, C.c_void_p(Y.data_ptr()), m, d,
            float(variance), float(lengthscale), C.c_void_p(out.data_ptr()), ld, 0, m,
            _abi.F64 if dtype == torch.float64 else _abi.F32,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tal_gram_stationary")
        return out[:, :n]

    def covariance(self, X, dtype: torch.dtype = torch.float64) -> torch.Tensor:
        return self.cross_covariance(X, X, dtype=dtype)


class Parameterized(Kernel):
    default_params: dict = {}

    def __init__(self, **params):
        self.params = {**self.__class__.default_params, **(params if params is not None else {})}
