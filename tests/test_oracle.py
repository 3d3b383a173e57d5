"""CPU tests of the oracle: the reference's known-answer values and the
mathematical identities that cross-check the restated functions."""
import json
import os

import numpy as np
import pytest

import oracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def test_hyperopt_nll0_gaussian_and_laplace():
    g = GOLD["hyperopt_nll0"]
    X = np.array(g["X"])
    y = np.sin(X).T[0]
    noise = np.full(5, g["noise_std"])
    nll_g = O.negative_log_likelihood(O.Gaussian, dict(O.Gaussian.default_params), X, y, noise)
    nll_l = O.negative_log_likelihood(O.Laplace, dict(O.Laplace.default_params), X, y, noise)
    assert nll_g == pytest.approx(g["gaussian_default_params_nll"], abs=g["abs_tol"])
    assert nll_l == pytest.approx(g["laplace_default_params_nll"], abs=g["abs_tol"])


def _prior_model(domain):
    kernel = O.Gaussian(lengthscale=1)
    distr = O.GaussianDistribution(mean=np.zeros(len(domain)), covariance=kernel.covariance(domain))
    return O.MarginalModel(key=0, domain=domain, distrs=[distr], beta=np.array([1.0]),
                           noise_rate=O.HomoscedasticNoise(1, np.array([0.1])),
                           use_objective_as_constraint=True)


def test_discrete_model_prior_state_exact():
    model = _prior_model(np.linspace(-1, 1, 100).reshape(-1, 1))
    assert np.all(model.means == 0.0)
    assert np.all(model.stddevs == 1.0)
    assert np.all(model.l == -1.0)
    assert np.all(model.u == 1.0)


def test_continuous_model_prior_state_exact():
    model = _prior_model(np.array(GOLD["continuous_model_prior"]["domain"]))
    assert np.all(model.means == 0.0) and np.all(model.stddevs == 1.0)
    assert np.all(model.l == -1.0) and np.all(model.u == 1.0)


def _small_model(N=60, q=2, seed=0, fc_obj=False):
    rng = np.random.default_rng(seed)
    domain = np.sort(rng.random((N, 2)), axis=0)
    kernels = [O.Matern52(lengthscale=0.4), O.Gaussian(lengthscale=0.3)][:q]
    distrs = [O.GaussianDistribution(np.zeros(N), k.covariance(domain)) for k in kernels]
    return O.MarginalModel(0, domain, distrs, beta=np.full(q, 2.0),
                           noise_rate=O.HomoscedasticNoise(q, np.full(q, 0.1)),
                           use_objective_as_constraint=fc_obj), rng


def test_batch_equals_sequential_conditioning():
    model, rng = _small_model()
    idx = np.array([3, 17, 41])
    y = rng.standard_normal((2, 3))
    noise = np.full((2, 3), 0.1)
    batch = [model.distr(i).posterior(idx, y[i], noise[i]) for i in range(2)]
    for t in range(3):
        model.step_idx(idx[t:t + 1], y[:, t:t + 1], noise[:, t:t + 1])
    for i in range(2):
        np.testing.assert_allclose(model.means[i], batch[i].mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(model.covariances[i], batch[i].covariance, rtol=1e-9, atol=1e-12)


def test_rank1_formula_matches_posterior():
    model, rng = _small_model(q=1)
    S = model.covariances[0].copy()
    mu = model.means[0].copy()
    x, yv, rho = 11, 0.7, 0.1
    model.step_idx(np.array([x]), np.array([[yv]]), np.array([[rho]]))
    k = S[x]
    s = k[x] + rho ** 2
    np.testing.assert_allclose(model.covariances[0], S - np.outer(k, k) / s, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(model.means[0], mu + k * (yv - mu[x]) / s, rtol=1e-12, atol=1e-15)


def test_directed_itl_forward_solve_identity():
    """diag(cov - K^T S^-1 K) == var - ||L^-1 K||^2 (what the CUDA path computes)."""
    model, rng = _small_model(N=80, q=1)
    for t in range(5):
        model.step_idx(np.array([rng.integers(80)]), rng.standard_normal((1, 1)), np.full((1, 1), 0.1))
    roi_idx = np.sort(rng.choice(80, 20, replace=False))
    F = O.itl_directed(model, roi_idx=roi_idx)
    cov = model.covariances[0]
    S = cov[np.ix_(roi_idx, roi_idx)] + 1e-10 * np.eye(20)
    L = np.linalg.cholesky(S)
    import scipy.linalg as sl
    V = sl.solve_triangular(L, cov[roi_idx, :], lower=True)
    pv = np.diag(cov) - np.sum(V * V, axis=0)
    F2 = 0.5 * np.maximum(np.log((np.diag(cov) + 0.01) / (pv + 0.01)), 0)
    np.testing.assert_allclose(F, F2, rtol=1e-7, atol=1e-9)


def test_directed_itl_whole_domain_vs_undirected():
    """Conditioning on the whole domain leaves ~no variance: F = 0.5*log(1 + var/rho^2)."""
    model, _ = _small_model(N=40, q=1)
    F = O.itl_directed(model, roi_idx=np.arange(40))
    np.testing.assert_allclose(F, 0.5 * O.itl_undirected(model), rtol=1e-5)


def test_vtl_closed_form_vs_bruteforce():
    model, rng = _small_model(N=30, q=2)
    roi_idx = np.array([2, 5, 9, 20])
    F = O.vtl_F(model, roi_idx=roi_idx)
    brute = np.zeros(30)
    nv = np.square(model.noise_rate.at(model.domain))
    for i in range(2):
        cov, var = model.covariances[i], model.variances[i]
        for x in range(30):
            brute[x] -= sum(var[a] - cov[a, x] ** 2 / (var[x] + nv[i, x]) for a in roi_idx)
    np.testing.assert_allclose(F, brute, rtol=1e-12)


def test_objective_argmax_checks_and_ties():
    model, _ = _small_model(N=20, q=1, fc_obj=True)
    model._l[0, :] = -1.0
    with pytest.raises(Exception, match="Operating safe set is empty"):
        model.objective_tie_set(np.zeros(20))
    model._l[0, 3:8] = 1.0
    with pytest.raises(Exception, match="negative infinity"):
        model.objective_tie_set(np.full(20, -np.inf))
    F = np.zeros(20)
    F[3:8] = np.nan
    with pytest.raises(Exception, match="NaN everywhere"):
        model.objective_tie_set(F)
    F = np.arange(20.0)
    F[5] = F[7] = 100.0
    F[10] = 1000.0  # unsafe, ignored
    idx, mx = model.objective_tie_set(F)
    assert list(idx) == [5, 7] and mx == 100.0 and model.objective_argmax_index(F) == 5


def test_is_subset_is_scalarwise():
    X = np.array([[0.0, 1.0]])
    A = np.array([[1.0, 0.0]])
    assert O.is_subset_of(X, A)  # row-wise it would be False (quirk 5)
    assert not O.is_subset_of(np.array([[0.0, 2.0]]), A)
    assert O.is_subset_of(np.zeros((0, 2)), np.zeros((0, 2)))
    assert not O.is_subset_of(np.zeros((0, 2)), A)
    assert not O.is_subset_of(X, np.zeros((0, 2)))


def test_cross_covariance_is_y_by_x():
    X = np.random.default_rng(0).random((5, 2))
    Y = np.random.default_rng(1).random((3, 2))
    K = O.Matern32(lengthscale=0.5).cross_covariance(X, Y)
    assert K.shape == (3, 5)
    tau = np.linalg.norm(X[4] - Y[1]) / 0.5
    assert K[1, 4] == pytest.approx((1 + np.sqrt(3) * tau) * np.exp(-np.sqrt(3) * tau))


def test_get_indices_first_min_and_mask():
    X = np.array([[0.0], [1.0], [1.0], [2.0]])
    assert list(O.get_indices(X, np.array([[1.0], [1.9]]))) == [1, 3]
    assert list(O.get_mask(X, X)) == [True, True, False, True]  # duplicate resolves to first copy


def test_step_uses_old_t_for_beta():
    N = 10
    domain = np.linspace(0, 1, N)[:, None]
    distr = O.GaussianDistribution(np.zeros(N), O.Gaussian().covariance(domain))
    seen = []
    def beta(t):
        seen.append(t)
        return np.array([1.0 + t])
    model = O.MarginalModel(0, domain, [distr], beta, O.HomoscedasticNoise(1, np.array([0.1])))
    seen.clear()
    model.step_idx(np.array([2]), np.array([[0.3]]), np.array([[0.1]]))
    assert set(seen) == {0} and model.t == 1


def test_low_rank_replay_matches_dense_restatement():
    """oracle/replay.py (the full-size checker: no N x N matrix) against the dense restatement of the
    reference at a size where both run: means / variances / bounds to the last bits (same operations in the same
    order), directed-ITL scores of a candidate subset to solver rounding."""
    from oracle.replay import LowRankReplay
    rng = np.random.default_rng(0)
    N, m, T = 300, 40, 12
    domain = rng.random((N, 3))
    A = np.sort(rng.choice(N, m, replace=False))
    kern = O.Matern52(lengthscale=0.3)
    noise = O.HomoscedasticNoise(1, np.array([0.1]))
    distrs = O.prior_distr(noise, domain, np.zeros((0, 3)), np.zeros((1, 0)), [kern], [O.ZeroMean()])
    model = O.MarginalModel(0, domain, distrs, beta=np.array([1.5]), noise_rate=noise)
    rep = LowRankReplay(kern, domain, 0.1, 1.5)
    C = np.sort(rng.choice(N, 50, replace=False))
    for t in range(T):
        F = O.itl_directed(model, roi_idx=A)
        np.testing.assert_allclose(rep.itl_directed(A, C), F[C], rtol=1e-6, atol=1e-9)
        x = int(np.argmax(F)) if t % 3 else int(rng.integers(N))
        y = float(rng.standard_normal())
        model.step_idx(np.array([x]), np.array([[y]]), np.array([[0.1]]))
        rep.observe(x, y)
        # (LAPACK's 1 x 1 solve scales by a reciprocal where the replay divides: last-bit differences)
        np.testing.assert_allclose(rep.var, model.variances[0], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(rep.mean, model.means[0], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(rep.l, model.l[0], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(rep.u, model.u[0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(rep.block(A[:7], C), model.covariances[0][np.ix_(A[:7], C)], rtol=1e-12, atol=1e-15)
