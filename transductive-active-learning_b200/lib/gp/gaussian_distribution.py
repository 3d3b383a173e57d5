"""Multivariate Gaussian over the finite domain (reference:
lib/gp/gaussian_distribution.py:13-121).

The covariance lives in a `talgp.DeviceGP` slot (HBM-resident, row stride
padded to 32 elements); `mean` / `covariance` / `variance` are views of that
state.  `posterior` runs the gather -> small Cholesky -> rank-m streaming
update kernels; like the reference it returns a NEW distribution unless
`inplace=True` (which avoids the second N x N buffer the reference always
pays for)."""
from __future__ import annotations

from typing import Optional

import torch

from lib._device import as_f64
from talgp.engine import DeviceGP

DEFAULT_JITTER = 1e-5  # gaussian_distribution.py:13


class GaussianDistribution:
    def __init__(self, mean, covariance, dtype: Optional[torch.dtype] = None, storage: str = "full"):
        mean = as_f64(mean).reshape(-1)
        n = mean.shape[0]
        cov = covariance if isinstance(covariance, torch.Tensor) else as_f64(covariance)
        dtype = dtype or (cov.dtype if cov.dtype in (torch.float32, torch.float64) else torch.float64)
        gp = DeviceGP(n, 1, dtype=dtype, device=mean.device, storage=storage)
        gp.load_covariance(0, cov)
        gp.mean[0].copy_(mean)
        self._gp, self._slot = gp, 0

    @classmethod
    def _from_engine(cls, gp: DeviceGP, slot: int) -> "GaussianDistribution":
        self = cls.__new__(cls)
        self._gp, self._slot = gp, slot
        return self

    @classmethod
    def initialize(cls, mean, kernel, X, dtype: torch.dtype = torch.float64, storage: str = "full"):
        """:69-71 — the Gram matrix is built in place on the device."""
        X = as_f64(X)
        gp = DeviceGP(X.shape[0], 1, dtype=dtype, device=X.device, storage=storage)
        variance, lengthscale = kernel._gram_params()
        gp.build_gram(0, kernel.kind, variance, lengthscale, X)
        gp.mean[0].copy_(mean.vector(X))
        return cls._from_engine(gp, 0)

    @property
    def n(self) -> int:
        return self._gp.N

    @property
    def mean(self) -> torch.Tensor:
        return self._gp.mean[self._slot]

    @property
    def covariance(self) -> torch.Tensor:
        return self._gp.covariance_view(self._slot)

    @property
    def variance(self) -> torch.Tensor:
        """:78-82 — maintained incrementally, bit-identical to diag(covariance)."""
        return self._gp.var[self._slot]

    @property
    def stddev(self) -> torch.Tensor:
        return torch.sqrt(self.variance)

    def sub_distr(self, idx) -> "GaussianDistribution":
        """:182-187 — the Gaussian limited to the points in `idx`."""
        idx = torch.as_tensor(idx, device=self._gp.device).to(torch.int64).reshape(-1)
        cov = self.covariance.index_select(0, idx).index_select(1, idx)
        return GaussianDistribution(self.mean.index_select(0, idx), cov, dtype=self._gp.dtype,
                                    storage=self._gp.storage)

    def sample(self, key, sample_shape) -> torch.Tensor:
        """:90-98 — independent functions from the Gaussian: mean + L z with L L^T = covariance from ONE
        blocked DMMA Cholesky (csrc/condvar.cu) and one pass over L for every 8 draws (csrc/dense.cu);
        the reference runs one SVD of the N x N matrix per draw.  A relative jitter of 1e-8 max(var) keeps the
        factorisation finite for numerically semi-definite posteriors.  The draws z come from an integer
        key (lib/_random.py), not threefry: distribution-equivalent only."""
        from lib import _random
        shape = tuple(sample_shape) if not isinstance(sample_shape, int) else (sample_shape,)
        n = self.n
        k = 1
        for s_ in shape:
            k *= int(s_)
        z = _random.normal(key, (k, n))
        return self._gp.sample(self._slot, z).reshape(shape + (n,))

    @property
    def total_variance(self) -> torch.Tensor:
        """:123-127."""
        return self.variance.sum()

    def entropy(self, jitter: float = 0.0) -> torch.Tensor:
        """:141-149 — 0.5 (n log(2 pi e) + logdet(covariance + jitter I)); the log-determinant is read off
        the Cholesky factor (NaN when the matrix is not numerically positive definite, where the
        reference's LU-based slogdet returns the log of |det|)."""
        import math
        n = self.n
        return 0.5 * (n * math.log(2 * math.pi * math.e) + self._gp.block_logdet(self._slot, None, n, None, float(jitter)))

    def information_gain(self, roi_idx, obs_idx, obs_noise_std=None, jitter: float = DEFAULT_JITTER) -> torch.Tensor:
        """:151-180 — mutual information of the points `roi_idx` and noisy observations at `obs_idx`."""
        gp = self._gp
        roi_idx = torch.as_tensor(roi_idx, device=gp.device).to(torch.int64).reshape(-1).contiguous()
        obs_idx = torch.as_tensor(obs_idx, device=gp.device).to(torch.int64).reshape(-1).contiguous()
        ns = None
        if obs_noise_std is not None:
            ns = as_f64(obs_noise_std).reshape(-1)
            if ns.numel() == 1 and obs_idx.numel() > 1:
                ns = ns.expand(obs_idx.numel()).contiguous()
        return gp.information_gain(self._slot, roi_idx, int(roi_idx.numel()), obs_idx, int(obs_idx.numel()), ns, jitter)

    def _clone(self) -> "GaussianDistribution":
        src = self._gp
        src.ensure_full()
        gp = DeviceGP(src.N, 1, dtype=src.dtype, device=src.device, storage=src.storage)
        gp.cov[0].copy_(src.cov[self._slot])
        gp.mean[0].copy_(src.mean[self._slot])
        gp.var[0].copy_(src.var[self._slot])
        return GaussianDistribution._from_engine(gp, 0)

    def posterior(self, obs_idx, obs, obs_noise_std=None, jitter: float = DEFAULT_JITTER,
                  inplace: bool = False) -> "GaussianDistribution":
        """:100-121 — Gaussian posterior given m observations `obs` at `obs_idx`
        with noise `obs_noise_std` (jitter^2 on the diagonal when None)."""
        out = self if inplace else self._clone()
        gp, i = out._gp, out._slot
        obs_idx = torch.as_tensor(obs_idx).reshape(-1)
        m = obs_idx.numel()
        if m == 0:  # :23-24
            return out
        obs = as_f64(obs).reshape(-1)
        if obs_noise_std is not None:
            ns = as_f64(obs_noise_std).reshape(-1)
            if ns.numel() == 1 and m > 1:
                ns = ns.expand(m).contiguous()
        else:
            ns = None
        gp.condition(i, obs_idx, obs, ns, jitter)
        return out


def compute_posterior_covariance(prior_covariance, obs_idx, obs_noise_std=None,
                                 jitter: float = DEFAULT_JITTER) -> torch.Tensor:
    """:40-55 — posterior covariance given observations at `obs_idx` (out of place)."""
    n = prior_covariance.shape[0]
    distr = GaussianDistribution(torch.zeros((n,), dtype=torch.float64), prior_covariance)
    m = torch.as_tensor(obs_idx).numel()
    distr.posterior(obs_idx, torch.zeros((m,), dtype=torch.float64), obs_noise_std, jitter,
                    inplace=True)
    return distr.covariance
