"""Non-stationary kernels (reference: lib/gp/kernels/non_stationary.py:7-20)."""
from lib.gp.kernels.base import Parameterized
from talgp import _abi


class Linear(Parameterized):
    kind = _abi.KERNEL_LINEAR
    default_params = {"variance": 1.0}

    def _gram_params(self):
        return float(self.params["variance"]), 1.0
