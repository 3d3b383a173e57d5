"""Mean-marginal ITL (reference: lib/algorithms/mm_itl.py:7-17)."""
import torch

from lib.algorithms import DirectedDiscreteGreedyAlgorithm
from talgp import _abi


class MMITL(DirectedDiscreteGreedyAlgorithm):
    def F(self) -> torch.Tensor:
        """sum over the ROI and the GPs of -0.5 log(1 - Cor^2)."""
        roi = self.roi
        return self.model._gp.score_rows(_abi.RULE_MMITL, roi.idx, roi.m, roi_mask=roi.mask)
