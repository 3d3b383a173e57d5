"""Line Bayesian optimisation (reference: lib/algorithms/line_bo.py:19-242):
adaptively selects one-dimensional subspaces, discretises them and runs a
nested discrete algorithm on the finite sub-domain.  The host logic mirrors the
reference; the per-line GP (ContinuousModel.marginalize) and the nested step run
on the device."""
from __future__ import annotations

from abc import abstractmethod
from typing import Callable, Optional

import numpy as np
import torch

from lib import _random
from lib._device import as_f64, device
from lib.algorithms import DirectedDiscreteGreedyAlgorithm, DiscreteGreedyAlgorithm, GreedyAlgorithm
from lib.function import Function
from lib.model.continuous import ContinuousModel
from lib.model.marginal import MarginalModel
from lib.utils import grad_approx

AlgConstructor = Callable[[MarginalModel], DiscreteGreedyAlgorithm]


def alg_constructor(Alg, **args) -> AlgConstructor:
    return lambda model: Alg(model=model, **args)


def directed_alg_constructor(Alg, roi_constructor) -> AlgConstructor:
    return lambda model: Alg(model=model, roi_constructor=roi_constructor)


class LineBO(GreedyAlgorithm):
    normalized_direction: Optional[torch.Tensor] = None

    def __init__(self, alg: AlgConstructor, model: ContinuousModel, bounds, n: int, eps: float, x_init=None):
        self.alg = alg
        self.model = model
        self.bounds = as_f64(bounds)
        self.n = n
        self.eps = eps
        self._x_init = as_f64(x_init) if x_init is not None else self.origin
        self.encountered_max = torch.empty((0, self.bounds.shape[0]), dtype=torch.float64, device=device())

    @property
    def d(self) -> int:
        return self.bounds.shape[0]

    @abstractmethod
    def _direction(self, f: Function) -> torch.Tensor:
        """A vector denoting the direction of the linear subspace."""

    def _normalized_direction(self, f: Function) -> torch.Tensor:
        """:80-86."""
        v = as_f64(self._direction(f))
        norm = torch.linalg.norm(v)
        if float(norm) < self.eps or bool(torch.isinf(norm)):
            return _coordinate_lbo(self)
        return v / norm

    @property
    def origin(self) -> torch.Tensor:
        return (self.bounds[:, 0] + self.bounds[:, 1]) / 2

    @property
    def x_init(self) -> torch.Tensor:
        if self.model.t == 0:
            return self._x_init
        return self.model.X[-1]

    def subspace_bounds(self):
        """:101-124."""
        assert self.normalized_direction is not None
        nd = self.normalized_direction
        disp = ((self.bounds - self.x_init[:, None]) / nd[:, None]).reshape(-1)
        pos, neg = disp[disp > 0], disp[disp <= 0]
        if pos.numel() == 0 or bool(torch.isinf(pos.min())):
            pos, neg = disp[disp >= 0], disp[disp < 0]
        return neg.max(), pos.min()

    def project(self, X) -> torch.Tensor:
        assert self.normalized_direction is not None
        return self.x_init + torch.outer(as_f64(X).reshape(-1), self.normalized_direction)

    def domain(self) -> torch.Tensor:
        """:130-138."""
        lo, hi = self.subspace_bounds()
        length = float(hi - lo)
        # jnp.linspace on the reference's fp64 bounds
        sub = torch.as_tensor(np.linspace(float(lo), float(hi), num=int(length * self.n)), device=device())
        return self.project(sub)

    def extended_domain(self) -> torch.Tensor:
        """:140-146 — previously observed points and previous guesses for the optimum included."""
        return torch.cat([self.model.X, self.domain(), self.encountered_max])

    def step(self, f: Function) -> dict:
        """:148-163."""
        self.normalized_direction = self._normalized_direction(f)
        domain = self.extended_domain()
        model = self.model.marginalize(domain)
        alg = self.alg(model)
        X, y, result = alg._step(f, update_model=False)
        self.model.step(X, y)
        argmax = model.argmax
        self.encountered_max = torch.cat([self.encountered_max, argmax.reshape(1, -1)])
        result["marginal_model"] = model
        result["argmax"] = argmax
        return result


class RandomLineBO(LineBO):
    def _direction(self, f: Function) -> torch.Tensor:
        return _random.uniform(self.acquire_key(), (self.d,))


class CoordinateLineBO(LineBO):
    def _direction(self, f: Function) -> torch.Tensor:
        return _coordinate_lbo(self)


class DescentLineBO(LineBO):
    """Direction from the gradient of f_0 (:183-231)."""

    def __init__(self, alg, model, bounds, n, eps, m: int, alpha: float, x_init=None):
        super().__init__(alg, model, bounds, n, eps, x_init)
        self.m = m
        self.alpha = alpha

    def _direction(self, f: Function) -> torch.Tensor:
        for _ in range(self.m):
            key = self.model.acquire_key()
            g = grad_approx(x=self.x_init,
                            f=lambda X: self.model.distr(X=X, i=0).sample(key=key, sample_shape=(1,))[0])
            if float(torch.linalg.norm(g)) < self.eps:
                return _coordinate_lbo(self)
            x = self.x_init - self.alpha * g
            X = torch.maximum(torch.minimum(x.reshape(1, -1), self.bounds[:, 1]), self.bounds[:, 0])
            y = f.observe(X)
            self.model.step(X=X, y=y)
        return grad_approx(x=self.x_init, f=lambda X: self.model.distr(X=X, i=0).mean)


def _coordinate_lbo(lbo: LineBO) -> torch.Tensor:
    i = _random.randint(lbo.acquire_key(), 0, lbo.d)
    return torch.eye(lbo.d, dtype=torch.float64, device=device())[i]
