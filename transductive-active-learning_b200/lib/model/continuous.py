"""GP model over a continuous domain (reference: lib/model/continuous.py:15-91):
finite multivariate Gaussians are recomputed on demand, conditioning on the
entire history of observations.

`distr(X, i)` is the reference's composition on the device: the Gram kernel
builds K over concat(X, X_train) in place (tal_gram_stationary), the t
observations are conditioned on with the general-m update kernels
(tal_gather_rows -> tal_small_posterior_weights -> tal_rankm_update, 64
observations per pass) and the leading n x n block is taken (sub_distr)."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from lib._device import as_f64, device
from lib.gp.gaussian_distribution import GaussianDistribution
from lib.gp.means import ZeroMean
from lib.model import Model
from lib.model.marginal import MarginalModel
from lib.noise import Noise


class ContinuousModel(Model):
    def __init__(self, key, d: int, prior_kernels: List, beta, noise_rate: Noise,
                 use_objective_as_constraint: bool = False, X=None, y=None,
                 prior_means: Optional[List] = None):
        q = len(prior_kernels)
        if not callable(beta):
            beta = np.asarray(beta.detach().cpu() if isinstance(beta, torch.Tensor) else beta,
                              dtype=np.float64).reshape(-1)
        super().__init__(key, q, beta, noise_rate, use_objective_as_constraint)
        self._prior_means = prior_means if prior_means is not None else [ZeroMean() for _ in range(q)]
        self._prior_kernels = prior_kernels
        dev = device()
        self.X = as_f64(X) if X is not None else torch.empty((0, d), dtype=torch.float64, device=dev)
        self.y = as_f64(y) if y is not None else torch.empty((q, 0), dtype=torch.float64, device=dev)

    def marginalize(self, X) -> MarginalModel:
        """:53-63."""
        X = as_f64(X)
        return MarginalModel(key=self.acquire_key(), domain=X, distrs=self.distrs(X), beta=self.beta,
                             noise_rate=self.noise_rate,
                             use_objective_as_constraint=self.first_constraint == 0, t=self.t)

    @property
    def t(self) -> int:
        return self.X.shape[0]

    def distr(self, X, i: int) -> GaussianDistribution:
        """:69-87 — distribution of f_i limited to the points in X."""
        X = as_f64(X)
        n, t = X.shape[0], self.t
        X_train, y_train = self.X[:t], self.y[i, :t]
        X_dom = torch.cat((X, X_train), dim=0).contiguous()
        prior = GaussianDistribution.initialize(mean=self._prior_means[i], kernel=self._prior_kernels[i], X=X_dom)
        if t > 0:
            prior.posterior(obs_idx=n + torch.arange(t, device=X.device), obs=y_train,
                            obs_noise_std=self.noise_rate.at(X_train)[i], inplace=True)
        return prior.sub_distr(torch.arange(n, device=X.device))

    def distrs(self, X) -> List[GaussianDistribution]:
        return [self.distr(X, i) for i in range(self.q)]

    def step(self, X, y):
        """:93-95."""
        X = as_f64(X).reshape(-1, self.X.shape[1])
        y = as_f64(y).reshape(self.q, -1)
        self.X = torch.cat((self.X, X), dim=0)
        self.y = torch.cat((self.y, y), dim=1)
