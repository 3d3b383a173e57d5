"""Correlation-based transductive learning (reference: lib/algorithms/ctl.py:7-17)."""
import torch

from lib.algorithms import DirectedDiscreteGreedyAlgorithm
from talgp import _abi


class CTL(DirectedDiscreteGreedyAlgorithm):
    def F(self) -> torch.Tensor:
        """sum over the ROI and the GPs of Cor[f_i(x'), y_i(x)]^2 (marginal.py:129-158)."""
        roi = self.roi
        return self.model._gp.score_rows(_abi.RULE_CTL, roi.idx, roi.m, roi_mask=roi.mask)
