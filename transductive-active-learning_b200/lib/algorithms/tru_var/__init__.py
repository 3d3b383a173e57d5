"""Truncated variance reduction (reference: lib/algorithms/tru_var/__init__.py:10-50)."""
import numpy as np
import torch

from lib.algorithms import DirectedDiscreteGreedyAlgorithm
from lib.algorithms.tru_var.model import TruVarModel
from talgp import _abi


class TruVar(DirectedDiscreteGreedyAlgorithm):
    def F(self) -> torch.Tensor:
        """-sum_{a in ROI} max(beta (var[a] - cov[a, x]^2 / (var[x] + rho[x]^2)), eta^2): the
        streaming row reduce with the TruVar element function (csrc/score.cu)."""
        assert self.model.q == 1 and self.model.first_constraint == 1
        roi = self.roi
        self.model.update_variance_threshold(roi)
        beta = float(np.asarray(self.model._beta_now())[0])
        return self.model._gp.score_rows(_abi.RULE_TRUVAR, roi.idx, roi.m, p0=[beta], p1=float(self.model.eta) ** 2, roi_mask=roi.mask)
