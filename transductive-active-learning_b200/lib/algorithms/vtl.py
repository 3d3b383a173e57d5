"""Variance-based transductive learning (reference: lib/algorithms/vtl.py:9-41)."""
import torch

from lib.algorithms import DirectedDiscreteGreedyAlgorithm
from talgp import _abi


class VTL(DirectedDiscreteGreedyAlgorithm):
    def F(self) -> torch.Tensor:
        """F[x] = -sum_i sum_{a in A} (var_i[a] - cov_i[a,x]^2 / (var_i[x] + rho_i[x]^2))."""
        roi = self.roi
        return self.model._gp.score_rows(_abi.RULE_VTL, roi.idx, roi.m, roi_mask=roi.mask)
