"""Finite-domain GP model (reference: lib/model/marginal.py:22-368, numeric
members).  Same constructor, properties and methods; the state lives in HBM
inside a `talgp.DeviceGP` and every update / mask / arg-max is a hand-written
sm_100a kernel behind include/tal_b200.h.

Differences that are documented rather than hidden:
  * arrays are `torch.Tensor`s on the GPU instead of `jax.Array`s;
  * `step` updates the covariance IN PLACE (the reference rebinds a new
    N x N array; none of the in-scope callers keep the old one);
  * exact ties in `objective_argmax` resolve to the LOWEST index (the
    reference draws uniformly with jax.random.choice, marginal.py:346-347);
    `last_argmax.n_ties` reports the size of the tie set.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import numpy as np
import torch

from lib._device import as_f64
from lib.gp.gaussian_distribution import GaussianDistribution
from lib.model import Model
from lib.noise import Noise
from lib.utils import get_indices, is_subset_of, set_to_idx
from talgp import _abi
from talgp.engine import DeviceGP

_MESSAGES = {
    _abi.ARGMAX_EMPTY_SAFE_SET: "Operating safe set is empty. Cannot make observations.",
    _abi.ARGMAX_ALL_NEG_INF: "Objective is negative infinity everywhere.",
    _abi.ARGMAX_ALL_NAN: "Objective is NaN everywhere.",
}


class ROI:
    """Region of interest: a subset of the model's finite domain.

    The reference passes ROIs around as point arrays [m, d] and re-derives
    their indices with a nearest-neighbour search on every use
    (lib/algorithms/__init__.py:101-135, lib/utils.py:53-56).  This object
    keeps the ascending domain indices (device) and the size next to the
    points so that no N x m x d search is needed; `points` materialises the
    reference's array on demand."""

    def __init__(self, model: "MarginalModel", idx: torch.Tensor, m: int,
                 mask: Optional[torch.Tensor] = None, fixed: bool = False):
        self.model = model
        self.idx = idx  # int64 device tensor, first m entries valid (ascending)
        self.m = int(m)
        self.mask = mask
        self.fixed = fixed  # the constructor returns this same object at every step

    @classmethod
    def from_mask(cls, model: "MarginalModel", mask: torch.Tensor) -> "ROI":
        m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
        idx, count = model._gp.mask_to_indices(m8)
        return cls(model, idx, int(count.item()), m8)

    @classmethod
    def from_points(cls, model: "MarginalModel", points) -> "ROI":
        points = as_f64(points).reshape(-1, model.d)
        idx = get_indices(model.domain, points) if points.shape[0] > 0 else torch.zeros(
            (0,), dtype=torch.int64, device=model.domain.device)
        return cls(model, idx, points.shape[0], None)

    @property
    def points(self) -> torch.Tensor:
        return self.model.domain[self.idx[: self.m]]

    @property
    def shape(self):
        return (self.m, self.model.d)

    def __len__(self) -> int:
        return self.m


class MarginalModel(Model):
    """lib/model/marginal.py:22-51."""

    def __init__(self, key, domain, distrs: List[GaussianDistribution], beta, noise_rate: Noise,
                 use_objective_as_constraint: bool = False, t: int = 0):
        q = len(distrs)
        if not callable(beta):
            beta = np.asarray(beta.detach().cpu() if isinstance(beta, torch.Tensor) else beta,
                              dtype=np.float64).reshape(-1)
        super().__init__(key, q, beta, noise_rate, use_objective_as_constraint)
        self.domain = as_f64(domain)
        n = self.domain.shape[0]
        gp0 = distrs[0]._gp
        shared = all(d._gp is gp0 and d._slot == i for i, d in enumerate(distrs)) and gp0.q == q
        if shared:
            self._gp = gp0  # adopt: built in one allocation by prior_distr
        else:  # jnp.array([distr.covariance ...]) (:47-48): one copy into a [q, n, ld] block
            gp = DeviceGP(n, q, dtype=gp0.dtype, device=self.domain.device, storage=gp0.storage)
            for i, d in enumerate(distrs):
                d._gp.ensure_full()
                gp.cov[i].copy_(d._gp.cov[d._slot])
                gp.mean[i].copy_(d._gp.mean[d._slot])
                gp.var[i].copy_(d._gp.var[d._slot])
            self._gp = gp
        self._gp.rho.copy_(self.noise_rate.at(self.domain))
        self._t = t
        self._gp.init_bounds(self._beta_now())  # _recompute_l / _recompute_u (:50-51)
        self._sid = None  # scalar ids for the isin-style subset test
        self.last_argmax = None

    # -- sizes -------------------------------------------------------------
    @property
    def n(self) -> int:
        return self.domain.shape[0]

    @property
    def d(self) -> int:
        return self.domain.shape[1]

    @property
    def t(self) -> int:
        return self._t

    def _beta_now(self) -> np.ndarray:
        b = self.beta(self._t)
        if isinstance(b, torch.Tensor):
            b = b.detach().cpu().numpy()
        return np.asarray(b, dtype=np.float64).reshape(-1)

    def get_indices(self, X) -> torch.Tensor:
        """:65-67."""
        if isinstance(X, ROI):
            return X.idx[: X.m]
        return get_indices(self.domain, X)

    def distr(self, i: int) -> GaussianDistribution:
        """:69-72 (a view of the model's state, not a copy)."""
        return GaussianDistribution._from_engine(self._gp, int(i))

    # -- state ---------------------------------------------------------------
    @property
    def l(self) -> torch.Tensor:
        return self._gp.l

    @property
    def u(self) -> torch.Tensor:
        return self._gp.u

    @property
    def means(self) -> torch.Tensor:
        return self._gp.mean

    @property
    def covariances(self) -> torch.Tensor:
        """[q, n, n] view of the HBM-resident covariances (row-sharded model:
        this rank's rows [q, n_loc, n], see `row_range`)."""
        return self._gp.covariances_view()

    @property
    def row_range(self):
        """(row0, n_loc): the covariance rows / candidates this rank holds."""
        return self._gp.row0, self._gp.n_loc

    @property
    def variances(self) -> torch.Tensor:
        """:103-106 — kept incrementally; equals vmap(diag)(covariances) bit for bit."""
        return self._gp.var

    @property
    def stddevs(self) -> torch.Tensor:
        """:108-111 (no clamping: a negative variance gives NaN)."""
        return torch.sqrt(self._gp.var)

    def masked_distr(self, i: int, mask) -> GaussianDistribution:
        """:74-81 — GP i limited to the points of `mask` (a copy of the block)."""
        idx = torch.nonzero(torch.as_tensor(mask, device=self._gp.device).reshape(-1).to(torch.bool))[:, 0]
        return self.distr(i).sub_distr(idx)

    def entropy(self, jitter: float = 0.0) -> torch.Tensor:
        """:113-117 — entropy, treating all functions as independent."""
        return sum(self.distr(i).entropy(jitter=jitter) for i in range(self.q))

    def masked_entropy(self, mask, jitter: float = 0.0) -> torch.Tensor:
        """:119-127 — entropy within `mask`: the masked block is gathered and factored in place of the
        reference's `masked_distr(...).entropy()` copy (csrc/dense.cu, csrc/condvar.cu)."""
        import math
        gp = self._gp
        idx = torch.nonzero(torch.as_tensor(mask, device=gp.device).reshape(-1).to(torch.bool))[:, 0].contiguous()
        m = int(idx.numel())
        total = torch.zeros((), dtype=torch.float64, device=gp.device)
        for i in range(self.q):
            total = total + 0.5 * (m * math.log(2 * math.pi * math.e) + gp.block_logdet(i, idx, m, None, float(jitter)))
        return total

    def sqd_correlation(self, X, A=None) -> torch.Tensor:
        """:129-158 — [q, m, n] squared correlations (small inputs; the CTL /
        MM-ITL rules use the streaming row-reduce kernels instead)."""
        if A is None:
            A = X
        X_idx = self.get_indices(X)
        A_idx = self.get_indices(A)
        if self._gp.sharded:  # the rows of X are assembled from the shards (collective), 256 at a time
            gp = self._gp
            cov = torch.stack([torch.cat([gp.gather_full_rows(i, X_idx[s: s + 256])[:, A_idx].clone()
                                          for s in range(0, X_idx.numel(), 256)]) for i in range(self.q)])
        else:
            cov = self.covariances[:, X_idx][:, :, A_idx].to(torch.float64)  # [q, n, m]
        sqd = torch.square(cov)
        noise = self.noise_rate.at(X.points if isinstance(X, ROI) else as_f64(X))
        var_X = self.variances[:, X_idx]
        var_A = self.variances[:, A_idx]
        out = sqd / ((var_X + torch.square(noise))[:, :, None] * var_A[:, None, :])
        return out.permute(0, 2, 1).contiguous()

    # -- safe sets (fused mask kernel, marginal.py:171-224) -------------------
    def _masks(self, want_pm=False, want_pe=False):
        op = None
        if type(self).operating_safe_set is not MarginalModel.operating_safe_set:
            op = self.operating_safe_set.view(torch.uint8)  # subclass override (unsafe.py:9-13)
        return self._gp.bounds_masks(self.first_constraint, op, want_pm, want_pe)

    @property
    def optimistic_safe_set(self) -> torch.Tensor:
        return self._masks()[1].view(torch.bool)

    @property
    def pessimistic_safe_set(self) -> torch.Tensor:
        return self._masks()[0].view(torch.bool)

    @property
    def operating_safe_set(self) -> torch.Tensor:
        """:181-184 — defaults to the pessimistic safe set."""
        return self.pessimistic_safe_set

    def _sid_table(self):
        if self._sid is None:
            dom = self.domain.detach().cpu().numpy()
            uniq, inv = np.unique(dom.reshape(-1), return_inverse=True)
            self._sid = (torch.as_tensor(inv.astype(np.int32), device=self.domain.device)
                         .reshape(dom.shape).contiguous(), int(uniq.shape[0]))
        return self._sid

    def safe_set_is_subset_of(self, A) -> bool:
        """:193-195 with is_subset_of's scalar-wise `isin` semantics (utils.py:65-71)."""
        if not isinstance(A, ROI):
            return is_subset_of(self.domain[self.operating_safe_set], A)
        # no constraint GP and the default operating safe set: the safe set is the whole domain at
        # every step (an `all` over an empty set, :176-184), so the answer for a fixed ROI never changes
        static = (self.first_constraint >= self.q and A.fixed
                  and type(self).operating_safe_set is MarginalModel.operating_safe_set)
        if static and getattr(A, "_covers_domain", None) is not None:
            return A._covers_domain
        import ctypes as C
        sid, n_sid = self._sid_table()
        gp = self._gp
        present = gp._scratch("subset_present", (n_sid,), torch.uint8)
        out = gp._scratch("subset_out", (1,), torch.int32)
        safe = self.operating_safe_set.view(torch.uint8)
        _abi.check(gp.lib.tal_scalar_subset(
            C.c_void_p(sid.data_ptr()), self.d, self.n, n_sid, C.c_void_p(safe.data_ptr()),
            C.c_void_p(A.idx.data_ptr()), None, A.m, C.c_void_p(present.data_ptr()),
            C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "tal_scalar_subset")
        res = bool(out.item())
        if static:
            A._covers_domain = res
        return res

    @property
    def potential_expanders(self) -> torch.Tensor:
        return self._masks(want_pe=True)[3].view(torch.bool)

    def argmax_within(self, mask) -> torch.Tensor:
        assert int(mask.sum()) > 0
        return self.domain[mask][torch.argmax(self.l[0, mask])]

    @property
    def argmax(self) -> torch.Tensor:
        return self.argmax_within(self.operating_safe_set)

    @property
    def max_l(self) -> torch.Tensor:
        return self._masks()[4][0]

    @property
    def potential_maximizers(self) -> torch.Tensor:
        return self._masks(want_pm=True)[2].view(torch.bool)

    @property
    def ucb(self) -> torch.Tensor:
        """:314-318."""
        safe = self.optimistic_safe_set
        return safe & (self.u[0] == torch.max(self.u[0, safe]))

    @property
    def max_u(self) -> torch.Tensor:
        safe = self.optimistic_safe_set
        vals = self.u[0, safe]
        return vals.max() if vals.numel() else torch.tensor(-np.inf, device=self.u.device)

    @property
    def pessimistic_max_u(self) -> torch.Tensor:
        safe = self.pessimistic_safe_set
        vals = self.u[0, safe]
        return vals.max() if vals.numel() else torch.tensor(-np.inf, device=self.u.device)

    # -- sampling (marginal.py:226-312) -----------------------------------------
    def sample(self, k: int, keys=None) -> torch.Tensor:
        """:226-236 — k independent functions from every GP: [q, k, n]."""
        if keys is None:
            keys = [self.acquire_key() for _ in range(self.q)]
        return torch.stack([self.distr(i).sample(key=keys[i], sample_shape=(k,)) for i in range(self.q)])

    def _ts_samples(self, k: int) -> torch.Tensor:
        """[k, q, n]: one factorisation per GP and one GEMM for all k samples (the
        reference factorises every GP again for every sample, :264-265)."""
        return self.sample(k).permute(1, 0, 2)

    @staticmethod
    def thompson_masks_from_samples(samples: torch.Tensor, I, J):
        """The per-sample logic of :264-273 on given samples [k, q, n]:
        (safety & optimality [k, n], safe optimum [k])."""
        I = torch.as_tensor(I, device=samples.device)
        J = torch.as_tensor(J, device=samples.device)
        n = samples.shape[2]
        cons = samples[:, J] if J.dtype == torch.bool else samples.index_select(1, J.to(torch.int64))
        obj = samples[:, I] if I.dtype == torch.bool else samples.index_select(1, I.to(torch.int64))
        safety = (cons >= 0).all(dim=1) if cons.shape[1] > 0 else torch.ones(
            (samples.shape[0], n), dtype=torch.bool, device=samples.device)
        masked = torch.where(safety[:, None, :], obj, torch.full_like(obj, -np.inf))
        safe_optimum = masked.min(dim=1).values.max(dim=1).values
        optimality = obj.min(dim=1).values >= safe_optimum[:, None]
        return safety & optimality, safe_optimum

    def thompson_sampling(self, k: int, I=None, J=None, strict: bool = False) -> torch.Tensor:
        """:238-278 — up to k points, each the safe maximum of an independent sample."""
        if I is None:
            I = torch.tensor([0])
        assert torch.as_tensor(I).shape[0] > 0
        if J is None:
            J = torch.zeros(self.q, dtype=torch.bool)
        masks, _ = self.thompson_masks_from_samples(self._ts_samples(k), I, J)
        result = masks.any(dim=0)
        if strict and int(result.sum()) == 0:
            raise Exception("Thompson Sampling: No sample has a safe point.")
        return result

    def thompson_sampling_values(self, k: int, I=None, J=None) -> torch.Tensor:
        """:280-312 — safe maxima of k independent samples."""
        if I is None:
            I = torch.tensor([0])
        if J is None:
            J = torch.zeros(self.q, dtype=torch.bool)
        return self.thompson_masks_from_samples(self._ts_samples(k), I, J)[1]

    # -- arg-max ---------------------------------------------------------------
    def _safe_override(self) -> Optional[torch.Tensor]:
        """The operating safe set as a mask when a subclass overrides it (unsafe.py:9-13,
        safe_opt/model.py:34-37); None = the pessimistic safe set, fused into the arg-max."""
        if type(self).operating_safe_set is not MarginalModel.operating_safe_set:
            return self.operating_safe_set.view(torch.uint8)
        return None

    def _check_argmax(self, res) -> int:
        """The reference's three error checks (:336-342) on an arg-max record."""
        self.last_argmax = res
        if res.status != _abi.ARGMAX_OK:
            raise Exception(_MESSAGES[res.status])
        self.acquire_key()  # the reference consumes one key per call (:346)
        return int(res.idx)

    def _step_index(self, idx_dev: torch.Tensor, y, auto_select: bool = False) -> None:
        """`step` for one observation at a domain index already on the device (:350-368)."""
        beta = self._beta_now()  # evaluated at the OLD t (:366-368)
        if isinstance(y, torch.Tensor) and y.is_cuda:
            yv = y.reshape(-1)[: self.q] if y.numel() == self.q else y[:, 0].contiguous()
        else:  # host observation: staged through the engine's pinned buffer, asynchronous H2D
            yv = np.asarray(y.cpu() if isinstance(y, torch.Tensor) else y, dtype=np.float64).reshape(self.q, -1)[:, 0]
        self._gp.observe(idx_dev, yv, beta, auto_select=auto_select)
        self._t += 1

    def _step_index_begin(self, idx_dev: torch.Tensor) -> None:
        """The part of `step` for one observation that does not depend on its value (:350-368: the observed
        row, the weights, the variances, the conditioning state and the queued covariance update): a host
        black box is asked for y while the device works on this."""
        self._gp.ctx_observe(idx_dev, 0, None, 1, False, self._beta_now())  # beta at the OLD t (:366-368)

    def _step_index_end(self, idx_dev: torch.Tensor, y) -> None:
        """... and the part that does: means and confidence bounds."""
        if isinstance(y, torch.Tensor) and y.is_cuda:
            yv = y.reshape(-1)[: self.q] if y.numel() == self.q else y[:, 0].contiguous()
        else:
            yv = np.asarray(y.cpu() if isinstance(y, torch.Tensor) else y, dtype=np.float64).reshape(self.q, -1)[:, 0]
        gp = self._gp
        gp.ctx_observe_y(idx_dev, 0, gp.stage_y(yv), 1, False)
        self._t += 1

    def domain_host(self) -> np.ndarray:
        """Host copy of the domain (made once): a host black box gets its input without a device round trip."""
        if getattr(self, "_domain_host", None) is None:
            self._domain_host = self.domain.detach().cpu().numpy()
        return self._domain_host

    def _argmax_index(self, F: torch.Tensor, sample_mask: Optional[torch.Tensor] = None):
        """Masked arg-max kernel + the reference's three error checks
        (:336-342).  F is overwritten with the sample-region-masked objective."""
        safe = self._safe_override()
        gp = self._gp
        if gp.sharded and F.numel() == self.n:  # a full-length objective: this rank's candidates
            F = F[gp.cand0: gp.cand0 + gp.n_cand]
        return self._check_argmax(gp.read_result(gp.argmax_candidates(F, safe, self.first_constraint, sample_mask)))

    def objective_argmax(self, F) -> torch.Tensor:
        """:330-348 — safe point maximising F (lowest index among exact ties)."""
        F = as_f64(F).clone()
        return self.domain[self._argmax_index(F)]

    # -- update ------------------------------------------------------------------
    def step(self, X, y):
        """:350-368."""
        X = as_f64(X).reshape(-1, self.d)
        m = X.shape[0]
        beta = self._beta_now()  # evaluated at the OLD t (:366-368)
        if m == 1:
            self._step_index(self.get_indices(X), y)  # device int64[1]
            return
        if m > 1:
            X_idx = self.get_indices(X)
            noise_rate = self.noise_rate.at(X)
            y = as_f64(y)
            for i in range(self.q):
                self._gp.condition(i, X_idx, y[i], noise_rate[i])
            b = torch.as_tensor(beta, device=self.domain.device)[:, None]
            sd = torch.sqrt(self._gp.var)
            self._gp.l.copy_(torch.maximum(self._gp.l, self._gp.mean - b * sd))
            self._gp.u.copy_(torch.minimum(self._gp.u, self._gp.mean + b * sd))
        self._t += m
