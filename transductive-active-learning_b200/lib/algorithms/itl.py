"""Information-based transductive learning (reference: lib/algorithms/itl.py:11-60)."""
import torch

from lib.algorithms import DirectedDiscreteGreedyAlgorithm
from lib.model.marginal import ROI, MarginalModel


class ITL(DirectedDiscreteGreedyAlgorithm):
    """conditioning: "recompute" re-runs gather -> Cholesky -> TRSM at every step
    like the reference; "incremental" keeps L^-1 cov[A, :] resident and downdates
    it per observation while the ROI object is unchanged (csrc/condinc.cu);
    "auto" (default) = incremental for ROI constructors that return the same
    ROI object at every step (fixed / index ROIs), recompute otherwise."""

    def __init__(self, model, roi_constructor, sample_region=None, conditioning: str = "auto"):
        super().__init__(model, roi_constructor, sample_region)
        assert conditioning in ("auto", "recompute", "incremental")
        self.conditioning = conditioning

    def F(self) -> torch.Tensor:
        """:20-25 — undirected if the safe set is (scalar-wise) inside the ROI."""
        roi = self.roi
        if self.model.safe_set_is_subset_of(A=roi):
            return _undirected(model=self.model)
        inc = self.conditioning == "incremental" or (self.conditioning == "auto" and roi.fixed)
        return _directed(model=self.model, roi=roi, incremental=inc)


def _undirected(model: MarginalModel) -> torch.Tensor:
    """:28-30 — sum_i log(1 + var_i / rho_i^2)."""
    gp = model._gp
    F = gp.score_itl_undirected()
    return F[gp.cand0: gp.cand0 + gp.n_cand] if gp.sharded else F


def _directed(model: MarginalModel, roi, incremental: bool = False) -> torch.Tensor:
    """:33-60 — sum_i 0.5 max(log((var_i + rho_i^2)/(var_{i|A} + rho_i^2)), 0)
    with var_{i|A} from gather -> Cholesky -> forward TRSM -> column sums of
    squares (the reference forms the full N x N posterior and takes its diagonal)."""
    if not isinstance(roi, ROI):
        roi = ROI.from_points(model, roi)
    return model._gp.score_itl_directed(roi.idx, roi.m, incremental=incremental, roi_key=roi)
