"""Models carrying TruVar's variance threshold (reference: lib/algorithms/tru_var/model.py:8-71)."""
import numpy as np
import torch

from lib.model.continuous import ContinuousModel
from lib.model.marginal import MarginalModel


class TruVarModel(MarginalModel):
    def __init__(self, eta: float, delta: float, r: float, **kwargs):
        super().__init__(**kwargs)
        self.eta = eta
        self.delta = delta
        self.r = r

    def update_variance_threshold(self, roi):
        """:28-33 — shrink eta once every target's beta^(1/2)-scaled stddev is within (1 + delta) eta."""
        idx = self.get_indices(roi)
        beta0 = float(np.asarray(self._beta_now())[0])
        if bool((np.sqrt(beta0) * self.stddevs[:, idx] <= (1 + self.delta) * self.eta).all()):
            self.eta = self.r * self.eta


class TruVarContinuousModel(ContinuousModel):
    def __init__(self, eta: float, delta: float, r: float, **kwargs):
        super().__init__(**kwargs)
        self.eta = eta
        self.delta = delta
        self.r = r

    def marginalize(self, X) -> TruVarModel:
        from lib._device import as_f64
        X = as_f64(X)
        return TruVarModel(eta=self.eta, delta=self.delta, r=self.r, key=self.acquire_key(), domain=X,
                           distrs=self.distrs(X), beta=self.beta, noise_rate=self.noise_rate,
                           use_objective_as_constraint=self.first_constraint == 0, t=self.t)
