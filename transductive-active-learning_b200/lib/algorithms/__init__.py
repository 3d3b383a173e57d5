"""Acquisition rules over a finite domain (reference: lib/algorithms/__init__.py:30-179).

One greedy step = score F -> safe-set-masked arg-max -> observe -> rank-1
posterior update, all on the device; the host only sees the 32-byte arg-max
record and hands the observation back."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Optional

import torch

from lib._device import as_f64
from lib.function import Function
from lib.model import Model
from lib.model.marginal import ROI, MarginalModel
from lib.utils import get_mask


class Algorithm(ABC):
    @abstractmethod
    def step(self, f: Function) -> dict:
        pass


class GreedyAlgorithm(Algorithm):
    def __init__(self, model: Model):
        self.model = model

    def acquire_key(self):
        return self.model.acquire_key()


class DiscreteGreedyAlgorithm(GreedyAlgorithm):
    """:53-98."""

    def __init__(self, model: MarginalModel, sample_region=None):
        self.model = model
        self.sample_region = sample_region if sample_region is not None else model.domain
        self._sample_mask = None

    @property
    def n(self) -> int:
        return self.model.domain.shape[0]

    @abstractmethod
    def F(self) -> torch.Tensor:
        pass

    def sample_mask(self) -> torch.Tensor:
        """get_mask(domain, sample_region) (:80).  Both are fixed for the
        lifetime of the algorithm, so the N x N' x d search runs once."""
        if self._sample_mask is None:
            region = self.sample_region.points if isinstance(self.sample_region, ROI) \
                else as_f64(self.sample_region)
            self._sample_mask = get_mask(self.model.domain, region).view(torch.uint8)
        return self._sample_mask

    def _step(self, f: Function, update_model: bool = True):
        F = self.F()  # row-sharded model: this rank's candidates only
        idx = self.model._argmax_index(F, self.sample_mask())  # masks F in place (:80-81)
        F = self.model._gp.gather_candidates(F)
        X = self.model.domain[idx: idx + 1]
        # (a black box that declares `host_input` gets the row of the model's host copy of the domain)
        y = f.observe(torch.from_numpy(self.model.domain_host()[idx: idx + 1]) if getattr(f, "host_input", False) else X)
        if update_model:
            self.model.step(X=X, y=y)
        if not isinstance(y, torch.Tensor):  # a host black box: keep its observation on the host
            import numpy as np
            y = torch.as_tensor(np.asarray(y, dtype=np.float64))
        return X, y, {"x": X[0], "y": y[:, 0], "F": F, "F_max": F[idx]}

    def step(self, f: Function) -> dict:
        X, y, result = self._step(f)
        return result


ROIConstructor = Callable[[MarginalModel], ROI]


def compute_roi_mask(domain, roi_description) -> torch.Tensor:
    """:138-154 — union over k boxes of the per-dimension bound tests."""
    domain = as_f64(domain)
    desc = as_f64(roi_description)
    lo, hi = desc[:, :, 0], desc[:, :, 1]  # [k, d]
    inside = (domain[None, :, :] >= lo[:, None, :]) & (domain[None, :, :] <= hi[:, None, :])
    return inside.all(dim=2).any(dim=0)


def compute_roi(domain, roi_description) -> torch.Tensor:
    return as_f64(domain)[compute_roi_mask(domain, roi_description)]


def fixed_roi_constructor(roi_description) -> ROIConstructor:
    cache = {}

    def ctor(model: MarginalModel) -> ROI:
        if id(model) not in cache:  # the description and the domain never change
            roi = ROI.from_mask(model, compute_roi_mask(model.domain, roi_description))
            roi.fixed = True
            cache[id(model)] = roi
        return cache[id(model)]

    return ctor


def index_roi_constructor(indices) -> ROIConstructor:
    """Fixed ROI given by domain indices (the transductive-learning setting,
    examples/experiment/transductive_learning.py:77-82 of the reference)."""
    cache = {}

    def ctor(model: MarginalModel) -> ROI:
        if id(model) not in cache:
            idx = torch.as_tensor(indices, dtype=torch.int64).to(model.domain.device).contiguous()
            cache[id(model)] = ROI(model, idx, idx.numel(), None, fixed=True)
        return cache[id(model)]

    return ctor


def f_roi_constructor(model: MarginalModel) -> ROI:
    idx = torch.arange(model.n, dtype=torch.int64, device=model.domain.device)
    return ROI(model, idx, model.n, None)


def pe_roi_constructor(model: MarginalModel) -> ROI:
    return ROI.from_mask(model, model.potential_expanders)


def pm_roi_constructor(model: MarginalModel) -> ROI:
    return ROI.from_mask(model, model.potential_maximizers)


def ts_roi_constructor(k: int) -> ROIConstructor:
    """:127-132 — Thompson-sampling ROI with the constraint GPs as J."""
    def ctor(model: MarginalModel) -> ROI:
        J = torch.arange(model.first_constraint, model.q)
        return ROI.from_mask(model, model.thompson_sampling(k=k, J=J))
    return ctor


def ucb_roi_constructor(model: MarginalModel) -> ROI:
    return ROI.from_mask(model, model.ucb)


class DirectedDiscreteGreedyAlgorithm(DiscreteGreedyAlgorithm):
    """:161-179."""

    def __init__(self, model: MarginalModel, roi_constructor: ROIConstructor, sample_region=None):
        super().__init__(model, sample_region)
        self.roi_constructor = roi_constructor

    @property
    def roi(self) -> ROI:
        roi = self.roi_constructor(self.model)
        if not isinstance(roi, ROI):  # a user constructor returning points [m, d]
            roi = ROI.from_points(self.model, roi)
        return roi
