"""GP priors conditioned on initial observations (reference: lib/gp/prior.py:34-66).

All q covariances are built directly in one HBM allocation [q, n, ld] by the
Gram kernel (no host N x N, no second copy when `MarginalModel` adopts them)."""
from __future__ import annotations

from typing import List

import torch

from lib._device import as_f64
from lib.gp.gaussian_distribution import GaussianDistribution
from lib.utils import get_indices
from talgp.engine import DeviceGP


def prior_distr(noise_rate, domain, X, y, kernels, means,
                dtype: torch.dtype = torch.float64, comm=None, storage: str = "full",
                arith=None) -> List[GaussianDistribution]:
    """lib/gp/prior.py:34-66.  comm: a talgp.dist communicator to row-shard the
    covariances over the ranks (every rank calls this with the same arguments).
    storage: "full" (reference layout) or "lower" (only the blocks on or below the
    diagonal are maintained; talgp.engine.DeviceGP).  arith: "two_rounding" (default) or "fma",
    the arithmetic of the covariance update (talgp.engine.DeviceGP)."""
    domain = as_f64(domain)
    X = as_f64(X).reshape(-1, domain.shape[1])
    y = as_f64(y)
    q = y.shape[0]
    indices = get_indices(domain, X) if X.shape[0] > 0 else torch.zeros(
        (0,), dtype=torch.int64, device=domain.device)
    noise_rates = noise_rate.at(X)
    gp = DeviceGP(domain.shape[0], q, dtype=dtype, device=domain.device, comm=comm, storage=storage, arith=arith)
    out = []
    for i in range(q):
        variance, lengthscale = kernels[i]._gram_params()
        gp.build_gram(i, kernels[i].kind, variance, lengthscale, domain)
        gp.mean[i].copy_(means[i].vector(domain))
        distr = GaussianDistribution._from_engine(gp, i)
        distr.posterior(obs_idx=indices, obs=y[i], obs_noise_std=noise_rates[i], inplace=True)
        out.append(distr)
    return out
