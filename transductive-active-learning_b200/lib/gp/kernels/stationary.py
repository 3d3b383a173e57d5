"""Stationary kernels (reference: lib/gp/kernels/stationary.py:13-55).  The
scalar formulas live in csrc/gram.cu (`kernel_value`)."""
from lib.gp.kernels.base import Parameterized
from talgp import _abi


class Stationary(Parameterized):
    default_params = {"variance": 1.0, "lengthscale": 1.0}

    def _gram_params(self):
        return float(self.params["variance"]), float(self.params["lengthscale"])


class Gaussian(Stationary):
    kind = _abi.KERNEL_GAUSSIAN


class Laplace(Stationary):
    kind = _abi.KERNEL_LAPLACE


class Matern32(Stationary):
    kind = _abi.KERNEL_MATERN32


class Matern52(Stationary):
    kind = _abi.KERNEL_MATERN52
