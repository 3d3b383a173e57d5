"""NumPy fp64 restatement of the reference's discrete-GP hot path (ORACLE).

TEST INFRASTRUCTURE: the checker for the CUDA path, never the product path.
Every function cites the reference file:line (relative to /root/reference) it
follows.  The arithmetic is restated *as written* (out-of-place N x N updates,
full N x N posterior inside directed ITL, general LU solves on the Cholesky
factors), quirks included; see SURVEY.md section 8c for the quirk list.

Third-party arithmetic: the reference executes inside jax/jaxlib 0.4.30 (XLA
CPU: LAPACK potrf/getrf, Eigen dot; requirements.txt:2-3), which is absent from
this image.  NumPy/SciPy call the same LAPACK routines (potrf, gesv) through
OpenBLAS; elementwise functions (exp, sqrt, log) are glibc/NumPy's.

Randomness: the reference threads threefry keys (jax.random) through
`acquire_key`; that PRNG is not reproducible without JAX.  The only place it
touches the path is the uniform choice among *exact* ties in
`objective_argmax`; here the full tie set is exposed and the default choice is
the lowest index (the documented policy of the CUDA path).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = [
    "TruVar",
    "TruVarModel",
    "ContinuousModel",
    "LineBO",
    "UnknownFunction",
    "thompson_masks_from_samples",
    "compute_strong_information_ratio",
    "compute_information_ratio",
    "DEFAULT_JITTER",
    "solve_linear_system",
    "noise_covariance_matrix",
    "get_indices",
    "get_mask",
    "is_subset_of",
    "set_to_idx",
    "Kernel",
    "Gaussian",
    "Laplace",
    "Matern32",
    "Matern52",
    "Linear",
    "ZeroMean",
    "ConstantMean",
    "HomoscedasticNoise",
    "HeteroscedasticNoise",
    "compute_posterior_covariance",
    "_compute_posterior_covariance",
    "GaussianDistribution",
    "prior_distr",
    "MarginalModel",
    "UnsafeModel",
    "DiscreteGreedyAlgorithm",
    "DirectedDiscreteGreedyAlgorithm",
    "ITL",
    "VTL",
    "CTL",
    "MMITL",
    "itl_undirected",
    "itl_directed",
    "vtl_F",
    "compute_roi_mask",
    "compute_roi",
    "fixed_roi_constructor",
    "f_roi_constructor",
    "pe_roi_constructor",
    "pm_roi_constructor",
    "ucb_roi_constructor",
    "negative_log_likelihood",
    "TableFunction",
]

DEFAULT_JITTER = 1e-5  # lib/gp/gaussian_distribution.py:13


# --------------------------------------------------------------------------- #
# lib/utils.py
# --------------------------------------------------------------------------- #
def _cholesky_nan(A: np.ndarray) -> np.ndarray:
    """jnp.linalg.cholesky returns NaNs on failure instead of raising."""
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return np.full_like(A, np.nan)


def solve_linear_system(A: np.ndarray, B: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """lib/utils.py:10-24.  L = chol(A); X = solve(L.T, solve(L, B)) with
    *general* (LU, partial pivoting) solves on the triangular factors."""
    L = _cholesky_nan(A)
    if np.isnan(L).any():
        return np.full(B.shape, np.nan), L
    X = np.linalg.solve(L.T, np.linalg.solve(L, B))
    return X, L


def noise_covariance_matrix(k: int, noise_std, default: float = 0) -> np.ndarray:
    """lib/utils.py:44-50.  identity(k) * square(noise_std); the default is
    squared too (jitter 1e-5 -> 1e-10 on the diagonal)."""
    if noise_std is None:
        noise_std = default
    return np.identity(k) * np.square(noise_std)


def get_indices(X: np.ndarray, A: np.ndarray, chunk: int = 256) -> np.ndarray:
    """lib/utils.py:53-56.  argmin_j ||X[j] - A[k]||_2, first-min tie rule.
    The reference materialises the [n, m, d] broadcast; here the m axis is
    chunked (result-identical) so that the oracle can run at larger N."""
    X = np.asarray(X, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    m = A.shape[0]
    out = np.empty((m,), dtype=np.int64)
    for s in range(0, m, chunk):
        e = min(m, s + chunk)
        distances = np.linalg.norm(X[:, None] - A[s:e], axis=2)
        out[s:e] = np.argmin(distances, axis=0)
    return out


def get_mask(X: np.ndarray, A: np.ndarray) -> np.ndarray:
    """lib/utils.py:59-62."""
    n = X.shape[0]
    mask = np.zeros((n,), dtype=bool)
    if A.shape[0] > 0:
        mask[get_indices(X, A)] = True
    return mask


def is_subset_of(X: np.ndarray, A: np.ndarray) -> bool:
    """lib/utils.py:65-71.  NB: jnp.isin works on *flattened scalars*, not rows."""
    if X.shape[0] == 0:
        return A.shape[0] == 0
    if A.shape[0] == 0:
        return False
    return bool(np.all(np.isin(X, A)))


def set_to_idx(arr: np.ndarray) -> np.ndarray:
    """lib/utils.py:74-75."""
    return np.where(arr)[0]


# --------------------------------------------------------------------------- #
# lib/gp/kernels/{base,stationary,non_stationary}.py
# --------------------------------------------------------------------------- #
class Kernel:
    """lib/gp/kernels/base.py:12-25.  cross_covariance(X, Y) is vmap over X
    inside, vmap over Y outside -> result is [len(Y), len(X)] (quirk 9)."""

    default_params: dict = {}

    def __init__(self, **params):
        self.params = {**self.__class__.default_params, **params}

    def pair(self, diff: np.ndarray) -> np.ndarray:  # diff = x - y, [..., d]
        raise NotImplementedError

    def cross_covariance(self, X: np.ndarray, Y: np.ndarray, chunk: int = 1024) -> np.ndarray:
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        out = np.empty((Y.shape[0], X.shape[0]), dtype=np.float64)
        for s in range(0, Y.shape[0], chunk):
            e = min(Y.shape[0], s + chunk)
            out[s:e] = self.pair(X[None, :, :] - Y[s:e, None, :])
        return out

    def covariance(self, X: np.ndarray) -> np.ndarray:
        return self.cross_covariance(X, X)


def _norm2(diff):  # jnp.linalg.norm(., ord=2) on a vector
    return np.sqrt(np.sum(np.square(diff), axis=-1))


class Gaussian(Kernel):
    """lib/gp/kernels/stationary.py:23-28 (sqrt then square, as written)."""

    default_params = {"variance": 1.0, "lengthscale": 1.0}

    def pair(self, diff):
        return self.params["variance"] * np.exp(
            -0.5 * np.square(_norm2(diff) / self.params["lengthscale"])
        )


class Laplace(Kernel):
    """lib/gp/kernels/stationary.py:31-35."""

    default_params = {"variance": 1.0, "lengthscale": 1.0}

    def pair(self, diff):
        return self.params["variance"] * np.exp(
            -np.sum(np.abs(diff), axis=-1) / self.params["lengthscale"]
        )


class Matern32(Kernel):
    """lib/gp/kernels/stationary.py:38-45."""

    default_params = {"variance": 1.0, "lengthscale": 1.0}

    def pair(self, diff):
        tau = _norm2(diff) / self.params["lengthscale"]
        return self.params["variance"] * (1.0 + np.sqrt(3.0) * tau) * np.exp(-np.sqrt(3.0) * tau)


class Matern52(Kernel):
    """lib/gp/kernels/stationary.py:48-55."""

    default_params = {"variance": 1.0, "lengthscale": 1.0}

    def pair(self, diff):
        tau = _norm2(diff) / self.params["lengthscale"]
        return (
            self.params["variance"]
            * (1.0 + np.sqrt(5.0) * tau + 5.0 / 3.0 * np.square(tau))
            * np.exp(-np.sqrt(5.0) * tau)
        )


class Linear(Kernel):
    """lib/gp/kernels/non_stationary.py:11-13.  variance * x.T @ y."""

    default_params = {"variance": 1.0}

    def cross_covariance(self, X, Y, chunk: int = 1024):
        X = np.asarray(X, dtype=np.float64)
        Y = np.asarray(Y, dtype=np.float64)
        return np.sum((self.params["variance"] * X)[None, :, :] * Y[:, None, :], axis=-1)


# --------------------------------------------------------------------------- #
# lib/gp/means.py, lib/noise.py
# --------------------------------------------------------------------------- #
class ZeroMean:
    """lib/gp/means.py:18-20."""

    def vector(self, X):
        return np.zeros((X.shape[0],), dtype=np.float64)


class ConstantMean:
    """lib/gp/means.py:23-29."""

    def __init__(self, value):
        self.value = value

    def vector(self, X):
        return np.full((X.shape[0],), float(self.value), dtype=np.float64)


class HomoscedasticNoise:
    """lib/noise.py:28-35."""

    def __init__(self, q: int, noise_rates):
        self.q = q
        self._noise_rates = np.asarray(noise_rates, dtype=np.float64)

    def at(self, X):
        n = X.shape[0]
        return np.tile(self._noise_rates, (n, 1)).T


class HeteroscedasticNoise:
    """lib/noise.py:38-46."""

    def __init__(self, noise_rates: List[Callable]):
        self.q = len(noise_rates)
        self._noise_rates = noise_rates

    def at(self, X):
        return np.array([self._noise_rates[i](X) for i in range(self.q)], dtype=np.float64)


# --------------------------------------------------------------------------- #
# lib/gp/gaussian_distribution.py
# --------------------------------------------------------------------------- #
def _compute_posterior_covariance(
    prior_covariance: np.ndarray,
    obs_idx: np.ndarray,
    obs_noise_std=None,
    jitter: float = DEFAULT_JITTER,
) -> Tuple[np.ndarray, np.ndarray]:
    """lib/gp/gaussian_distribution.py:16-37 (out of place, as written)."""
    obs_idx = np.asarray(obs_idx, dtype=np.int64)
    if obs_idx.shape[0] == 0:  # :23-24
        return prior_covariance, np.array([])
    n = prior_covariance.shape[0]
    full_idx = np.arange(n)
    obs_noise_mat = noise_covariance_matrix(
        k=obs_idx.shape[0], noise_std=obs_noise_std, default=jitter
    )
    Sigma = prior_covariance[np.ix_(obs_idx, obs_idx)] + obs_noise_mat
    Kxt = prior_covariance[np.ix_(obs_idx, full_idx)]
    Sigma_inv_Kxt, _ = solve_linear_system(Sigma, Kxt)
    covariance = prior_covariance - Kxt.T @ Sigma_inv_Kxt
    return covariance, Sigma_inv_Kxt


def compute_posterior_covariance(prior_covariance, obs_idx, obs_noise_std=None, jitter=DEFAULT_JITTER):
    """lib/gp/gaussian_distribution.py:40-55."""
    return _compute_posterior_covariance(prior_covariance, obs_idx, obs_noise_std, jitter)[0]


class GaussianDistribution:
    """lib/gp/gaussian_distribution.py:58-187 (the on-path members)."""

    def __init__(self, mean: np.ndarray, covariance: np.ndarray):
        self.mean = mean
        self.covariance = covariance

    @classmethod
    def initialize(cls, mean, kernel, X):  # :69-71
        return cls(mean=mean.vector(X), covariance=kernel.covariance(X))

    @property
    def n(self) -> int:
        return self.mean.shape[0]

    @property
    def variance(self):  # :78-82
        return np.diagonal(self.covariance)

    @property
    def stddev(self):  # :84-88
        return np.sqrt(self.variance)

    def posterior(self, obs_idx, obs, obs_noise_std=None, jitter=DEFAULT_JITTER):  # :100-121
        obs_idx = np.asarray(obs_idx, dtype=np.int64)
        covariance, Sigma_inv_Kxt = _compute_posterior_covariance(
            self.covariance, obs_idx, obs_noise_std, jitter
        )
        if obs_idx.shape[0] == 0:
            return GaussianDistribution(mean=self.mean, covariance=covariance)
        mean = self.mean + Sigma_inv_Kxt.T @ (np.asarray(obs, dtype=np.float64) - self.mean[obs_idx])
        return GaussianDistribution(mean=mean, covariance=covariance)

    def log_prob(self, y):  # :129-138
        delta = y - self.mean
        alpha, L = solve_linear_system(self.covariance, delta)
        return -(
            0.5 * delta.T @ alpha
            + np.sum(np.log(np.diag(L)))
            + 0.5 * self.n * np.log(2 * np.pi)
        )

    def sub_distr(self, idx):  # :182-187
        idx = np.asarray(idx, dtype=np.int64)
        return GaussianDistribution(mean=self.mean[idx], covariance=self.covariance[np.ix_(idx, idx)])

    def entropy(self, jitter: float = 0.0):  # :141-149
        return 0.5 * (
            self.n * np.log(2 * np.pi * np.e)
            + np.linalg.slogdet(self.covariance + jitter * np.eye(self.covariance.shape[0]))[1]
        )

    def information_gain(self, roi_idx, obs_idx, obs_noise_std=None, jitter=DEFAULT_JITTER):  # :151-180
        roi_idx = np.asarray(roi_idx, dtype=np.int64)
        obs_idx = np.asarray(obs_idx, dtype=np.int64)
        prior_covariance = self.covariance[np.ix_(obs_idx, obs_idx)]
        posterior_covariance = compute_posterior_covariance(
            prior_covariance=self.covariance, obs_idx=roi_idx, obs_noise_std=None, jitter=jitter
        )[np.ix_(obs_idx, obs_idx)]
        obs_noise_mat = noise_covariance_matrix(k=obs_idx.shape[0], noise_std=obs_noise_std)
        predictive_covariance = prior_covariance + obs_noise_mat
        posterior_predictive_covariance = posterior_covariance + obs_noise_mat
        predictive_covariance_log_det = np.linalg.slogdet(predictive_covariance)[1]
        posterior_predictive_covariance_log_det = np.nan_to_num(
            np.linalg.slogdet(posterior_predictive_covariance)[1], nan=-np.inf
        )
        return 0.5 * np.maximum(predictive_covariance_log_det - posterior_predictive_covariance_log_det, 0)


def negative_log_likelihood(K, theta: dict, X, y, noise_std, mean=None):
    """lib/gp/hyperopt.py:14-36 — used only to pin the oracle against the
    reference's known-answer NLL_0 values (tests/test_hyperopt.py:29,45)."""
    n = X.shape[0]
    if mean is None:
        mean = np.zeros(n)
    K_theta = K(**theta).covariance(X) + noise_covariance_matrix(n, noise_std)
    return -GaussianDistribution(mean=mean, covariance=K_theta).log_prob(y)


# --------------------------------------------------------------------------- #
# lib/gp/prior.py
# --------------------------------------------------------------------------- #
def prior_distr(noise_rate, domain, X, y, kernels, means) -> List[GaussianDistribution]:
    """lib/gp/prior.py:34-66."""
    q = y.shape[0]
    indices = get_indices(domain, X) if X.shape[0] > 0 else np.zeros((0,), dtype=np.int64)
    noise_rates = noise_rate.at(X)
    out = []
    for i in range(q):
        covariance = kernels[i].covariance(domain)
        prior = GaussianDistribution(mean=means[i].vector(domain), covariance=covariance)
        out.append(prior.posterior(obs_idx=indices, obs=y[i], obs_noise_std=noise_rates[i]))
    return out


# --------------------------------------------------------------------------- #
# lib/model/__init__.py, lib/model/marginal.py, lib/model/unsafe.py
# --------------------------------------------------------------------------- #
class MarginalModel:
    """lib/model/marginal.py:22-368 (numeric members on the path)."""

    def __init__(self, key, domain, distrs, beta, noise_rate,
                 use_objective_as_constraint: bool = False, t: int = 0):
        # lib/model/__init__.py:26-45
        self._key = key
        self.q = len(distrs)
        self.first_constraint = 0 if use_objective_as_constraint else 1
        self.beta = (lambda t: beta) if not callable(beta) else beta
        self.noise_rate = noise_rate
        # lib/model/marginal.py:43-51
        self.domain = np.asarray(domain, dtype=np.float64)
        self._means = np.array([d.mean for d in distrs], dtype=np.float64)
        self._covariances = np.array([d.covariance for d in distrs], dtype=np.float64)
        self._t = t
        self._l = self._recompute_l()
        self._u = self._recompute_u()

    # -- sizes ---------------------------------------------------------------
    @property
    def n(self):
        return self.domain.shape[0]

    @property
    def d(self):
        return self.domain.shape[1]

    @property
    def t(self):
        return self._t

    def get_indices(self, X):  # :65-67
        return get_indices(self.domain, X)

    def distr(self, i):  # :69-72
        return GaussianDistribution(mean=self._means[i], covariance=self._covariances[i])

    def masked_distr(self, i, mask):  # :74-81
        mask_idx = set_to_idx(mask)
        return GaussianDistribution(mean=self._means[i, mask],
                                    covariance=self._covariances[i][np.ix_(mask_idx, mask_idx)])

    def entropy(self, jitter: float = 0.0):  # :113-117
        return sum(self.distr(i).entropy(jitter=jitter) for i in range(self.q))

    def masked_entropy(self, mask, jitter: float = 0.0):  # :119-127
        return sum(self.masked_distr(i, mask).entropy(jitter=jitter) for i in range(self.q))

    # -- state ---------------------------------------------------------------
    @property
    def l(self):
        return self._l

    @property
    def u(self):
        return self._u

    @property
    def means(self):
        return self._means

    @property
    def covariances(self):
        return self._covariances

    @property
    def variances(self):  # :103-106
        return np.array([np.diag(c) for c in self._covariances])

    @property
    def stddevs(self):  # :108-111 (no clamping: negative variance -> NaN)
        with np.errstate(invalid="ignore"):
            return np.sqrt(self.variances)

    def sqd_correlation(self, X, A=None):  # :129-158
        if A is None:
            A = X
        n = X.shape[0]
        X_idx = self.get_indices(X)
        A_idx = self.get_indices(A)
        obs_noise_var = np.square(self.noise_rate.at(X))
        variances = self.variances
        out = np.empty((self.q, A_idx.shape[0], n))
        for i in range(self.q):
            sqd_cross = np.square(self._covariances[i][np.ix_(X_idx, A_idx)])  # [n, m]
            var_X = variances[i, X_idx]
            var_A = variances[i, A_idx]
            for k in range(A_idx.shape[0]):
                out[i, k] = sqd_cross[:, k] / ((var_X + obs_noise_var[i]) * var_A[k])
        return out

    def _beta_col(self):
        return np.expand_dims(np.asarray(self.beta(self.t), dtype=np.float64), axis=1)

    def _recompute_l(self):  # :160-161
        return self.means - self._beta_col() * self.stddevs

    def _recompute_u(self):  # :163-164
        return self.means + self._beta_col() * self.stddevs

    # -- safe sets -----------------------------------------------------------
    @property
    def optimistic_safe_set(self):  # :171-174
        return np.all(self.u[self.first_constraint:] >= 0, axis=0)

    @property
    def pessimistic_safe_set(self):  # :176-179
        return np.all(self.l[self.first_constraint:] >= 0, axis=0)

    @property
    def operating_safe_set(self):  # :181-184
        return self.pessimistic_safe_set

    def safe_set_is_subset_of(self, A):  # :193-195
        return is_subset_of(self.domain[self.operating_safe_set], A)

    @property
    def potential_expanders(self):  # :197-200
        return self.optimistic_safe_set & ~self.pessimistic_safe_set

    @property
    def max_l(self):  # :217-219
        return np.max(self.l[0, self.operating_safe_set], initial=-np.inf)

    @property
    def potential_maximizers(self):  # :221-224
        return self.optimistic_safe_set & (self.u[0] >= self.max_l)

    @property
    def ucb(self):  # :314-318
        safe = self.optimistic_safe_set
        return safe & (self.u[0] == np.max(self.u[0, safe]))

    @property
    def max_u(self):  # :320-323
        return np.max(self.u[0, self.optimistic_safe_set], initial=-np.inf)

    @property
    def argmax(self):  # :202-210
        mask = self.operating_safe_set
        assert np.sum(mask) > 0
        return self.domain[mask][np.argmax(self.l[0, mask])]

    # -- argmax --------------------------------------------------------------
    def objective_tie_set(self, F):
        """marginal.py:330-344 up to the PRNG draw: returns (indices, safe_max)."""
        safe_set = self.operating_safe_set
        if np.sum(safe_set) == 0:
            raise Exception("Operating safe set is empty. Cannot make observations.")
        if np.all(F == -np.inf):
            raise Exception("Objective is negative infinity everywhere.")
        safe_F = F.copy()
        safe_F[~safe_set] = -np.inf
        if np.all(np.isnan(safe_F[safe_set])):
            raise Exception("Objective is NaN everywhere.")
        safe_max = np.nanmax(safe_F)
        indices = np.argwhere(safe_F == safe_max).reshape(-1)
        return indices, safe_max

    def objective_argmax_index(self, F) -> int:
        """Tie policy of this build: lowest index among exact ties (the
        reference draws uniformly with jax.random.choice, marginal.py:346-347)."""
        indices, _ = self.objective_tie_set(F)
        return int(indices[0])

    def objective_argmax(self, F):  # :330-348
        return self.domain[self.objective_argmax_index(F)]

    # -- update --------------------------------------------------------------
    def step(self, X, y):  # :350-368
        m = X.shape[0]
        X_idx = self.get_indices(X) if m > 0 else np.zeros((0,), dtype=np.int64)
        noise_rate = self.noise_rate.at(X)
        self.step_idx(X_idx, y, noise_rate)

    def step_idx(self, X_idx, y, noise_rate):
        """Same as `step` with the index lookup done by the caller."""
        m = len(X_idx)
        posts = [
            self.distr(i).posterior(obs_idx=X_idx, obs=y[i], obs_noise_std=noise_rate[i])
            for i in range(self.q)
        ]
        self._means = np.array([p.mean for p in posts])
        self._covariances = np.array([p.covariance for p in posts])
        # beta is evaluated at the OLD t (quirk 3)
        self._l = np.maximum(self._l, self._recompute_l())
        self._u = np.minimum(self._u, self._recompute_u())
        self._t += m


class UnsafeModel(MarginalModel):
    """lib/model/unsafe.py:9-13."""

    @property
    def operating_safe_set(self):
        return np.ones((self.n,), dtype=bool)


# --------------------------------------------------------------------------- #
# lib/algorithms/__init__.py
# --------------------------------------------------------------------------- #
def compute_roi_mask(domain, roi_description):
    """lib/algorithms/__init__.py:138-158."""
    D = domain.shape[1]
    k = roi_description.shape[0]
    res = np.zeros((domain.shape[0],), dtype=bool)
    for l in range(k):
        box = np.ones((domain.shape[0],), dtype=bool)
        for d in range(D):
            box &= (domain[:, d] >= roi_description[l, d, 0]) & (domain[:, d] <= roi_description[l, d, 1])
        res |= box
    return res


def compute_roi(domain, roi_description):
    """lib/algorithms/__init__.py:157-158."""
    return domain[compute_roi_mask(domain, roi_description)]


def fixed_roi_constructor(roi_description):  # :114-115
    return lambda model: compute_roi(model.domain, roi_description)


f_roi_constructor = lambda model: model.domain  # :118
pe_roi_constructor = lambda model: model.domain[model.potential_expanders]  # :119-121
pm_roi_constructor = lambda model: model.domain[model.potential_maximizers]  # :122-124
ucb_roi_constructor = lambda model: model.domain[model.ucb]  # :135


class DiscreteGreedyAlgorithm:
    """lib/algorithms/__init__.py:53-98."""

    def __init__(self, model, sample_region=None):
        self.model = model
        self.sample_region = sample_region if sample_region is not None else model.domain

    @property
    def n(self):
        return self.model.domain.shape[0]

    def F(self):
        raise NotImplementedError

    def masked_F(self):
        F = np.array(self.F(), dtype=np.float64, copy=True)
        F[~get_mask(self.model.domain, self.sample_region)] = -np.inf  # :80
        return F

    def _step(self, f, update_model: bool = True):
        F = self.masked_F()
        X = self.model.objective_argmax(F=F).reshape(1, -1)
        y = f.observe(X)
        if update_model:
            self.model.step(X=X, y=y)
        return X, y, {"x": X[0], "y": y[:, 0], "F": F, "F_max": F[self.model.get_indices(X)][0]}

    def step(self, f):
        return self._step(f)[2]


class DirectedDiscreteGreedyAlgorithm(DiscreteGreedyAlgorithm):
    """lib/algorithms/__init__.py:161-179."""

    def __init__(self, model, roi_constructor, sample_region=None):
        super().__init__(model, sample_region)
        self.roi_constructor = roi_constructor

    @property
    def roi(self):
        return self.roi_constructor(self.model)


# --------------------------------------------------------------------------- #
# lib/algorithms/itl.py
# --------------------------------------------------------------------------- #
def itl_undirected(model) -> np.ndarray:
    """lib/algorithms/itl.py:28-30 (no 1/2 factor, quirk 10)."""
    obs_noise_var = np.square(model.noise_rate.at(model.domain))
    return np.sum(np.log(1 + model.variances / obs_noise_var), axis=0)


def itl_directed(model, roi=None, roi_idx=None) -> np.ndarray:
    """lib/algorithms/itl.py:33-60: a full N x N posterior conditioned on the
    ROI (jitter^2 = 1e-10 on the diagonal), of which only the diagonal is used."""
    if roi_idx is None:
        roi_idx = model.get_indices(roi)
    covariances = model.covariances
    obs_noise_stds = model.noise_rate.at(model.domain)
    total = np.zeros((model.n,), dtype=np.float64)
    for i in range(model.q):
        prior_variance = np.diag(covariances[i])
        posterior_covariance = compute_posterior_covariance(
            prior_covariance=covariances[i], obs_idx=roi_idx, obs_noise_std=None
        )
        posterior_variance = np.diag(posterior_covariance)
        obs_noise = np.square(obs_noise_stds[i])  # diag(noise_covariance_matrix(...))
        predictive_variance = prior_variance + obs_noise
        posterior_predictive_variance = posterior_variance + obs_noise
        with np.errstate(invalid="ignore", divide="ignore"):
            total = total + 0.5 * np.maximum(
                np.log(predictive_variance / posterior_predictive_variance), 0
            )
    return total


class ITL(DirectedDiscreteGreedyAlgorithm):
    """lib/algorithms/itl.py:11-25."""

    def F(self):
        roi = self.roi
        if self.model.safe_set_is_subset_of(A=roi):
            return itl_undirected(self.model)
        return itl_directed(self.model, roi=roi)


# --------------------------------------------------------------------------- #
# lib/algorithms/vtl.py, ctl.py, mm_itl.py
# --------------------------------------------------------------------------- #
def vtl_F(model, roi=None, roi_idx=None) -> np.ndarray:
    """lib/algorithms/vtl.py:16-41:
    F[x] = - sum_i sum_{a in A} (var_i[a] - cov_i[a, x]^2 / (var_i[x] + rho_i[x]^2))."""
    if roi_idx is None:
        roi_idx = model.get_indices(roi)
    covariances = model.covariances
    noise_var = np.square(model.noise_rate.at(model.domain))
    total = np.zeros((model.n,), dtype=np.float64)
    for i in range(model.q):
        covariance = covariances[i]
        variance = np.diag(covariance)
        C = covariance[roi_idx, :]  # covariance[idx, obs_idx], idx in roi
        terms = variance[roi_idx][:, None] - C ** 2 / (variance + noise_var[i])[None, :]
        total = total + (-np.sum(terms, axis=0))
    return total


class VTL(DirectedDiscreteGreedyAlgorithm):
    """lib/algorithms/vtl.py:9-41."""

    def F(self):
        return vtl_F(self.model, roi=self.roi)


class CTL(DirectedDiscreteGreedyAlgorithm):
    """lib/algorithms/ctl.py:7-17."""

    def F(self):
        return np.sum(self.model.sqd_correlation(X=self.model.domain, A=self.roi), axis=(0, 1))


class MMITL(DirectedDiscreteGreedyAlgorithm):
    """lib/algorithms/mm_itl.py:7-17."""

    def F(self):
        sqd_cor = self.model.sqd_correlation(X=self.model.domain, A=self.roi)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sum(-0.5 * np.log(1 - sqd_cor), axis=(0, 1))


# --------------------------------------------------------------------------- #
# a deterministic stand-in for lib/function (the black box; harness-side)
# --------------------------------------------------------------------------- #
class TableFunction:
    """`Function.observe` (lib/function/__init__.py:5-16) backed by a table of
    pre-drawn noisy observations y_table[q, N] over the domain, so that the
    oracle and the CUDA path see identical observations."""

    def __init__(self, domain, y_table):
        self.domain = np.asarray(domain, dtype=np.float64)
        self.y_table = np.asarray(y_table, dtype=np.float64)
        self.q = self.y_table.shape[0]

    def observe(self, X):
        idx = get_indices(self.domain, np.asarray(X, dtype=np.float64))
        return self.y_table[:, idx]


# --------------------------------------------------------------------------- #
# continuous-domain model and LineBO (lib/model/continuous.py,
# lib/algorithms/line_bo.py); 8f-4: Thompson-sampling masks (marginal.py:238-312)
# --------------------------------------------------------------------------- #
class UnknownFunction:
    """lib/function/unknown.py:8-28."""

    def __init__(self, q, f, use_objective_as_constraint: bool = False):
        self.q = q
        self.first_constraint = 0 if use_objective_as_constraint else 1
        self._f = f

    def observe(self, X):
        X = np.asarray(X, dtype=np.float64)
        return np.array([np.asarray(self._f(X[j]), dtype=np.float64).reshape(-1) for j in range(X.shape[0])]).T


class ContinuousModel:
    """lib/model/continuous.py:15-95."""

    def __init__(self, key, d, prior_kernels, beta, noise_rate, use_objective_as_constraint=False,
                 X=None, y=None, prior_means=None):
        q = len(prior_kernels)
        self._key = key
        self.q = q
        self.first_constraint = 0 if use_objective_as_constraint else 1
        self.beta = (lambda t: beta) if not callable(beta) else beta
        self.noise_rate = noise_rate
        self._prior_means = prior_means if prior_means is not None else [ZeroMean() for _ in range(q)]
        self._prior_kernels = prior_kernels
        self.X = X if X is not None else np.empty((0, d))
        self.y = y if y is not None else np.empty((q, 0))

    def acquire_key(self):
        self._key = (int(self._key) * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return self._key

    def marginalize(self, X):  # :53-63
        return MarginalModel(key=self.acquire_key(), domain=X, distrs=self.distrs(X), beta=self.beta,
                             noise_rate=self.noise_rate,
                             use_objective_as_constraint=self.first_constraint == 0, t=self.t)

    @property
    def t(self):
        return self.X.shape[0]

    def distr(self, X, i):  # :69-87
        X = np.asarray(X, dtype=np.float64)
        n, t = X.shape[0], self.t
        X_train, y_train = self.X[:t], self.y[i, :t]
        X_dom = np.concatenate((X, X_train), axis=0)
        prior = GaussianDistribution.initialize(mean=self._prior_means[i], kernel=self._prior_kernels[i], X=X_dom)
        posterior = prior.posterior(obs_idx=n + np.arange(t), obs=y_train,
                                    obs_noise_std=self.noise_rate.at(X_train)[i])
        return posterior.sub_distr(np.arange(n))

    def distrs(self, X):
        return [self.distr(X, i) for i in range(self.q)]

    def step(self, X, y):  # :93-95
        self.X = np.concatenate((self.X, np.asarray(X, dtype=np.float64)), axis=0)
        self.y = np.concatenate((self.y, np.asarray(y, dtype=np.float64)), axis=1)


class LineBO:
    """lib/algorithms/line_bo.py:31-163 with the direction supplied by `_direction`."""

    normalized_direction = None

    def __init__(self, alg, model, bounds, n, eps, x_init=None):
        self.alg, self.model, self.bounds, self.n, self.eps = alg, model, np.asarray(bounds, dtype=np.float64), n, eps
        self._x_init = x_init if x_init is not None else self.origin
        self.encountered_max = np.empty((0, self.bounds.shape[0]))

    @property
    def d(self):
        return self.bounds.shape[0]

    def _direction(self, f):
        raise NotImplementedError

    def _normalized_direction(self, f):  # :80-86 (the fallback draws a coordinate direction)
        v = np.asarray(self._direction(f), dtype=np.float64)
        norm = np.linalg.norm(v)
        assert norm >= self.eps and not np.isinf(norm), "oracle: coordinate fallback needs a PRNG"
        return v / norm

    @property
    def origin(self):
        return (self.bounds[:, 0] + self.bounds[:, 1]) / 2

    @property
    def x_init(self):
        return self._x_init if self.model.t == 0 else self.model.X[-1]

    def subspace_bounds(self):  # :101-124
        with np.errstate(divide="ignore", invalid="ignore"):
            disp = ((self.bounds - self.x_init[:, None]) / self.normalized_direction[:, None]).reshape(-1)
        pos, neg = disp[disp > 0], disp[disp <= 0]
        if np.isinf(np.min(pos, initial=np.inf)):
            pos, neg = disp[disp >= 0], disp[disp < 0]
        return np.max(neg), np.min(pos)

    def project(self, X):
        return self.x_init + np.outer(X, self.normalized_direction)

    def domain(self):  # :130-138
        lo, hi = self.subspace_bounds()
        return self.project(np.linspace(lo, hi, num=int((hi - lo) * self.n)))

    def extended_domain(self):  # :140-146
        return np.concatenate([self.model.X, self.domain(), self.encountered_max])

    def step(self, f):  # :148-163
        self.normalized_direction = self._normalized_direction(f)
        domain = self.extended_domain()
        model = self.model.marginalize(domain)
        alg = self.alg(model)
        X, y, result = alg._step(f, update_model=False)
        self.model.step(X, y)
        argmax = model.argmax
        self.encountered_max = np.concatenate([self.encountered_max, argmax.reshape(1, -1)])
        result["marginal_model"] = model
        result["argmax"] = argmax
        return result


# lib/metrics/information_ratio.py:10-57
def compute_strong_information_ratio(F_max, prev=None):
    if len(F_max) < 1:
        return np.array(np.inf)
    F_max_latest = F_max[-1]
    new = np.nanmin(np.array(F_max) / F_max_latest)
    return np.minimum(new, prev) if prev is not None and not np.isnan(prev) else new


def compute_information_ratio(model, roi_idx, sample_region_idx):
    if roi_idx is None:
        roi_idx = np.arange(model.n)
    if sample_region_idx is None:
        sample_region_idx = np.arange(model.n)
    boundary_idx = sample_region_idx
    joint_info_gain = 0
    noise_rates = model.noise_rate.at(model.domain)
    for i in range(model.q):
        joint_info_gain += model.distr(i).information_gain(
            roi_idx=roi_idx, obs_idx=boundary_idx, obs_noise_std=noise_rates[i])
    indiv_info_gains = itl_directed(model, roi_idx=np.asarray(roi_idx))
    return np.sum(indiv_info_gains[boundary_idx]) / joint_info_gain


def thompson_masks_from_samples(samples, I, J):
    """The per-sample logic of MarginalModel.thompson_sampling / _values (marginal.py:264-273,
    304-310) on given samples [k, q, n]: (safety & optimality [k, n], safe optimum [k])."""
    samples = np.asarray(samples, dtype=np.float64)
    masks, values = [], []
    for s in samples:
        safety = np.all(s[J] >= 0, axis=0)
        safe_optimum = np.max(np.min(np.where(safety, s[I], -np.inf), axis=0))
        optimality = np.min(s[I], axis=0) >= np.repeat(safe_optimum, s.shape[1])
        masks.append(safety & optimality)
        values.append(safe_optimum)
    return np.array(masks), np.array(values)


# --------------------------------------------------------------------------- #
# TruVar (lib/algorithms/tru_var/__init__.py:10-50, model.py:8-33)
# --------------------------------------------------------------------------- #
class TruVarModel(MarginalModel):
    def __init__(self, eta, delta, r, **kwargs):
        super().__init__(**kwargs)
        self.eta, self.delta, self.r = eta, delta, r

    def update_variance_threshold(self, roi):  # model.py:28-33
        idx = self.get_indices(roi)
        if np.all(np.sqrt(self.beta(self.t)[0]) * self.stddevs[:, idx] <= (1 + self.delta) * self.eta):
            self.eta = self.r * self.eta


class TruVar(DirectedDiscreteGreedyAlgorithm):
    def F(self):  # __init__.py:17-50
        assert self.model.q == 1 and self.model.first_constraint == 1
        roi = self.roi
        self.model.update_variance_threshold(roi)
        beta = self.model.beta(self.model.t)[0]
        covariance = self.model.covariances[0]
        variance = self.model.variances[0]
        noise_var = np.square(self.model.noise_rate.at(self.model.domain)[0])
        idx = get_indices(self.model.domain, roi)
        post = variance[idx][:, None] - covariance[idx, :] ** 2 / (variance + noise_var)[None, :]  # [m, n]
        return -np.sum(np.maximum(beta * post, self.model.eta ** 2), axis=0)
