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

    gather_F = True  # row-sharded model: result["F"] is all-gathered to full length (one NCCL call per step)

    def _incremental(self, roi) -> bool:
        return self.conditioning == "incremental" or (self.conditioning == "auto" and roi.fixed)

    def _step(self, f, update_model: bool = True):
        """The greedy step (lib/algorithms/__init__.py:77-94).  With a resident conditioning state
        for a fixed ROI the selection of this step was already made by the fused kernels of the
        previous observation (csrc/fused.cu): the host reads the 32-byte record, asks the black box,
        and issues ONE call for the update + the next selection."""
        model, gp = self.model, self.model._gp
        st = gp._cond
        if update_model and st is not None and st["valid"] and st["append"] and gp._ctx_usable():
            roi = self.roi
            if st["key"] is roi and self._incremental(roi) and not model.safe_set_is_subset_of(A=roi):
                rec = gp.ctx_select(model.first_constraint, model._safe_override(), self.sample_mask())
                # a device-resident black box (a lookup: no host round trip) gets the whole update as one fused
                # call; any other black box is asked for y WHILE the device does everything that does not need
                # it (observed row, weights, variances, conditioning pass: ~all of the step): that part is
                # enqueued straight after the record has been read -- its index is taken from the record on the
                # device, and a failed arg-max is a no-op there -- then the means / bounds and the selection of
                # the next candidate are enqueued behind it
                if getattr(f, "device_resident", False):
                    idx = model._check_argmax(gp.read_result(rec))
                    F = gp._scratch("ctx_F", (gp.n_cand,)).clone()  # (the update overwrites it with the next scores)
                    X = model.domain[idx: idx + 1]
                    y = f.observe(X)
                    model._step_index(rec[0:1], y, auto_select=True)
                else:
                    c, stream = gp.ctx_observe_prepare(model._beta_now())  # beta at the OLD t (:366-368)
                    idx_ptr = rec.data_ptr()
                    res = gp.read_result(rec)                 # the one host synchronisation of the step
                    gp.ctx_observe_fire(c, stream, idx_ptr)   # device busy again
                    idx = model._check_argmax(res)
                    F = gp._scratch("ctx_F", (gp.n_cand,)).clone()  # (still this step's scores: rewritten by the select below)
                    if getattr(f, "host_input", False):
                        X = torch.from_numpy(model.domain_host()[idx: idx + 1])
                    else:
                        X = model.domain[idx: idx + 1]
                    y = f.observe(X)
                    model._step_index_end(rec[0:1], y)
                    if gp._cond is not None and gp._cond["valid"]:
                        gp.ctx_select(model.first_constraint, model._safe_override(), self.sample_mask())
                if gp.sharded and self.gather_F:
                    F = gp.gather_candidates(F)
                if not isinstance(y, torch.Tensor):
                    import numpy as np
                    y = torch.as_tensor(np.asarray(y, dtype=np.float64))
                i_loc = idx - (gp.cand0 if (gp.sharded and F.numel() != model.n) else 0)
                F_max = F[i_loc] if 0 <= i_loc < F.numel() else torch.as_tensor(float(model.last_argmax.val))
                return X, y, {"x": X[0], "y": y[:, 0], "F": F, "F_max": F_max}
        return super()._step(f, update_model)

    def F(self) -> torch.Tensor:
        """:20-25 — undirected if the safe set is (scalar-wise) inside the ROI."""
        roi = self.roi
        if self.model.safe_set_is_subset_of(A=roi):
            return _undirected(model=self.model)
        return _directed(model=self.model, roi=roi, incremental=self._incremental(roi))


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
