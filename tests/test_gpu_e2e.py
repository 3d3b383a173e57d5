"""GPU parity tests through the drop-in `lib.*` surface: whole BO / active-
learning trajectories against the oracle, free-running and teacher-forced,
on scaled-down versions of BASELINE.json's configs, plus size-independent
properties at the full N = 65,536."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

TIE_TOL = 1e-9  # near-tie tolerance of the selection policy


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def _bumps(x):  # an objective with several bumps on [-2, 8]
    return (np.exp(-(x - 1) ** 2) + 1.5 * np.exp(-(x - 5) ** 2 / 0.5) + 0.8 * np.exp(-(x - 6.5) ** 2 / 0.2)
            - 0.3 - 0.05 * (x - 3) ** 2 * (x < 0))


def _build_pair(domain, kernels_o, kernels_c, beta, rho, X0, y0, fc_obj, model_cls="MarginalModel"):
    from lib.gp.means import ZeroMean
    from lib.gp.prior import prior_distr
    from lib.model.marginal import MarginalModel
    from lib.model.unsafe import UnsafeModel
    from lib.noise import HomoscedasticNoise
    q = len(kernels_o)
    no = O.HomoscedasticNoise(q, np.asarray(rho))
    nc = HomoscedasticNoise(q, np.asarray(rho))
    do = O.prior_distr(no, domain, X0, y0, kernels_o, [O.ZeroMean()] * q)
    dc = prior_distr(nc, domain, X0, y0, kernels_c, [ZeroMean()] * q)
    Mo = O.UnsafeModel if model_cls == "UnsafeModel" else O.MarginalModel
    Mc = UnsafeModel if model_cls == "UnsafeModel" else MarginalModel
    mo = Mo(0, domain, do, beta=np.asarray(beta, float), noise_rate=no, use_objective_as_constraint=fc_obj)
    mc = Mc(0, domain, dc, beta=np.asarray(beta, float), noise_rate=nc, use_objective_as_constraint=fc_obj)
    return mo, mc


def _check_state(mo, mc, rtol=1e-9):
    n = mo.n
    np.testing.assert_allclose(_np(mc.means), mo.means, rtol=rtol, atol=1e-11)
    np.testing.assert_allclose(_np(mc.variances), mo.variances, rtol=rtol, atol=1e-11)
    np.testing.assert_allclose(_np(mc.l), mo.l, rtol=rtol, atol=1e-10)
    np.testing.assert_allclose(_np(mc.u), mo.u, rtol=rtol, atol=1e-10)


def _selection_ok(F_oracle_masked, safe, idx):
    sf = np.where(safe, F_oracle_masked, -np.inf)
    mx = np.nanmax(sf)
    return sf[idx] >= mx - TIE_TOL * max(1.0, abs(mx))


class _Table(O.TableFunction):
    pass


def _run_pair(alg_o, alg_c, table, T, teacher_forced, rtol_F=1e-8):
    """Runs T steps on both sides.  Teacher-forced: the CUDA side's selection is
    checked against the oracle's scores, then both sides apply the ORACLE's
    selection.  Free-running: each side applies its own selection and the
    trajectories must coincide up to ties within tolerance."""
    from lib.function import TableFunction
    mo, mc = alg_o.model, alg_c.model
    fo = O.TableFunction(mo.domain, table)
    fc = TableFunction(mc.domain, table, first_constraint=mc.first_constraint)
    agree = 0
    for t in range(T):
        Fo = alg_o.masked_F()
        Fc = alg_c.F()
        idx_c = mc._argmax_index(Fc, alg_c.sample_mask())
        fin = np.isfinite(Fo)
        np.testing.assert_allclose(_np(Fc)[fin], Fo[fin], rtol=rtol_F, atol=1e-9)
        assert np.array_equal(np.isneginf(_np(Fc)), np.isneginf(Fo))
        idx_o = mo.objective_argmax_index(Fo)
        assert _selection_ok(Fo, mo.operating_safe_set, idx_c), (t, idx_c, idx_o)
        agree += int(idx_c == idx_o)
        ties, _ = mo.objective_tie_set(Fo)
        assert mc.last_argmax.n_ties >= 1
        use_c = idx_o if teacher_forced else idx_c
        Xo = mo.domain[idx_o: idx_o + 1]
        mo.step(Xo, fo.observe(Xo))
        Xc = mc.domain[use_c: use_c + 1]
        mc.step(Xc, fc.observe(Xc))
        if teacher_forced or idx_c == idx_o:
            _check_state(mo, mc)
        else:
            return agree, t  # diverged at a documented near-tie: stop comparing states
    return agree, T


def test_reference_discrete_model_prior_state(cuda):
    """tests/test_discrete_model.py:29-39 of the reference, through the drop-in API."""
    from lib.gp.gaussian_distribution import GaussianDistribution
    from lib.gp import kernels
    from lib.model.marginal import MarginalModel
    from lib.noise import HomoscedasticNoise
    domain = torch.linspace(-1, 1, 100, dtype=torch.float64).reshape(-1, 1)
    N = domain.shape[0]
    kernel = kernels.stationary.Gaussian(lengthscale=1)
    distr = GaussianDistribution(mean=torch.zeros(N), covariance=kernel.covariance(domain))
    model = MarginalModel(key=0, domain=domain, distrs=[distr], beta=np.array([1.0]),
                          noise_rate=HomoscedasticNoise(q=1, noise_rates=np.array([0.1])),
                          use_objective_as_constraint=True)
    assert torch.all(model.means == torch.zeros((1, N), device=cuda, dtype=torch.float64))
    assert torch.all(model.stddevs == 1.0)
    assert torch.all(model.l == -1.0) and torch.all(model.u == 1.0)
    assert model.covariances.shape == (1, N, N) and model.n == N and model.d == 1 and model.t == 0


@pytest.mark.parametrize("teacher_forced", [True, False])
def test_c1_1d_safe_bo_itl_pm(cuda, teacher_forced):
    """Config 1 scaled down: 1-D safe BO, ITL with the potential-maximizer ROI,
    RBF, objective used as constraint (examples/.../1d_hard/experiment.py)."""
    from lib.algorithms import pm_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.gp.kernels import stationary
    N = 400
    domain = np.linspace(-2, 8, N)[:, None]
    rng = np.random.default_rng(8)
    rho = [0.1]
    table = (_bumps(domain[:, 0]) + 0.1 * rng.standard_normal(N))[None, :]
    seed_idx = np.sort(rng.choice(np.where((domain[:, 0] > 4.5) & (domain[:, 0] < 6))[0], 8))
    X0, y0 = domain[seed_idx], table[:, seed_idx]
    mo, mc = _build_pair(domain, [O.Gaussian(lengthscale=1.0)], [stationary.Gaussian(lengthscale=1.0)],
                         beta=[3.0], rho=rho, X0=X0, y0=y0, fc_obj=True)
    _check_state(mo, mc)
    alg_o = O.ITL(mo, O.pm_roi_constructor)
    alg_c = ITL(mc, pm_roi_constructor)
    agree, steps = _run_pair(alg_o, alg_c, table, T=25, teacher_forced=teacher_forced, rtol_F=1e-6)
    assert steps >= 5 and agree >= steps - 2


@pytest.mark.parametrize("rule", ["VTL-PM", "VTL-PE"])
def test_c2_2d_safe_bo_vtl(cuda, rule):
    """Config 2 scaled down: 2-D safe BO, objective + 1 constraint GP, VTL
    (examples/unknown_constraints/bo/2d_island/experiment.py)."""
    from lib.algorithms import pe_roi_constructor, pm_roi_constructor
    from lib.algorithms.vtl import VTL
    from lib.gp.kernels import stationary
    g = np.arange(-3, 3, 0.25)
    domain = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)  # 24 x 24
    N = len(domain)
    rng = np.random.default_rng(3)
    f = 2.0 * np.exp(-0.5 * np.sum((domain - 0.5) ** 2, axis=1))
    gfun = 1.5 * np.exp(-0.25 * np.sum(domain ** 2, axis=1)) - 0.4
    table = np.stack([f, gfun]) + 0.1 * rng.standard_normal((2, N))
    seed = np.array([int(np.argmin(np.sum(domain ** 2, axis=1)))])
    mo, mc = _build_pair(domain, [O.Gaussian(lengthscale=1.0)] * 2, [stationary.Gaussian(lengthscale=1.0)] * 2,
                         beta=[3.0, 3.0], rho=[0.1, 0.1], X0=domain[seed], y0=table[:, seed], fc_obj=False)
    ro = O.pm_roi_constructor if rule == "VTL-PM" else O.pe_roi_constructor
    rc = pm_roi_constructor if rule == "VTL-PM" else pe_roi_constructor
    alg_o, alg_c = O.VTL(mo, ro), VTL(mc, rc)
    agree, steps = _run_pair(alg_o, alg_c, table, T=15, teacher_forced=True)
    assert steps == 15 and agree >= 13


@pytest.mark.parametrize("conditioning", ["recompute", "incremental"])
def test_c3_fixed_roi_itl_matern52_d4(cuda, conditioning):
    """Config 3 scaled down: transductive ITL with a fixed target set, d = 4
    Matérn-5/2, unconstrained (always the directed rule), default sample region."""
    from lib.algorithms import index_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.gp.kernels import stationary
    g = (np.arange(6) + 0.0) / 6
    domain = np.stack(np.meshgrid(g, g, g, g, indexing="ij"), -1).reshape(-1, 4)  # 6^4 = 1296
    N = len(domain)
    rng = np.random.default_rng(0)
    # one uniform point per grid cell: on the bare grid every coordinate value
    # of the domain occurs in the target set and the reference's scalar-wise
    # subset test (lib/utils.py:65-71) would select the UNdirected rule
    domain = domain + rng.random(domain.shape) / 6
    A = np.sort(rng.choice(N, 160, replace=False))
    table = (np.sin(3 * domain.sum(axis=1)) + 0.1 * np.random.default_rng(1).standard_normal(N))[None, :]
    mo, mc = _build_pair(domain, [O.Matern52(lengthscale=0.25)], [stationary.Matern52(lengthscale=0.25)],
                         beta=[1.0], rho=[0.1], X0=np.zeros((0, 4)), y0=np.zeros((1, 0)), fc_obj=False)
    alg_o = O.ITL(mo, lambda model: model.domain[A])
    alg_c = ITL(mc, index_roi_constructor(A), conditioning=conditioning)
    agree, steps = _run_pair(alg_o, alg_c, table, T=12, teacher_forced=True, rtol_F=1e-6)
    assert steps == 12 and agree >= 10


def test_grid_domain_takes_undirected_branch_like_reference(cuda):
    """Quirk 5: on a bare grid `safe_set_is_subset_of` is true although the
    target set is a strict subset, so ITL.F is the undirected score on both sides."""
    from lib.algorithms import index_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.gp.kernels import stationary
    g = np.arange(5) / 5.0
    domain = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    A = np.array([0, 6, 12, 18, 24])  # the diagonal: contains every coordinate value
    mo, mc = _build_pair(domain, [O.Matern52(lengthscale=0.5)], [stationary.Matern52(lengthscale=0.5)],
                         beta=[1.0], rho=[0.1], X0=np.zeros((0, 2)), y0=np.zeros((1, 0)), fc_obj=False)
    alg_o = O.ITL(mo, lambda model: model.domain[A])
    alg_c = ITL(mc, index_roi_constructor(A))
    assert mo.safe_set_is_subset_of(mo.domain[A]) and mc.safe_set_is_subset_of(alg_c.roi)
    np.testing.assert_allclose(_np(alg_c.F()), O.itl_undirected(mo), rtol=1e-12)
    np.testing.assert_allclose(_np(alg_c.F()), alg_o.F(), rtol=1e-12)


def test_step_result_dict_and_errors(cuda):
    from lib.algorithms import f_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.function import TableFunction
    from lib.gp.kernels import stationary
    N = 120
    domain = np.linspace(0, 1, N)[:, None]
    table = np.sin(6 * domain[:, 0])[None, :]
    mo, mc = _build_pair(domain, [O.Gaussian(lengthscale=0.2)], [stationary.Gaussian(lengthscale=0.2)],
                         beta=[2.0], rho=[0.1], X0=np.zeros((0, 1)), y0=np.zeros((1, 0)), fc_obj=False,
                         model_cls="UnsafeModel")
    alg = ITL(mc, f_roi_constructor)  # ROI = domain -> undirected; N-way tie at t = 0 (quirk 11)
    res = alg.step(TableFunction(mc.domain, table))
    assert set(res) == {"x", "y", "F", "F_max"}
    assert res["x"].shape == (1,) and res["y"].shape == (1,) and res["F"].shape == (N,)
    assert mc.last_argmax.n_ties == N and mc.last_argmax.idx == 0  # lowest index among exact ties
    assert mc.t == 1 and float(res["F_max"]) == float(res["F"][0])
    # the reference's error messages (marginal.py:337,339,342)
    from lib.model.marginal import MarginalModel
    from lib.gp.gaussian_distribution import GaussianDistribution
    from lib.noise import HomoscedasticNoise
    m = MarginalModel(0, domain, [GaussianDistribution(np.full(N, -5.0), np.eye(N))], beta=np.array([1.0]),
                      noise_rate=HomoscedasticNoise(1, [0.1]), use_objective_as_constraint=True)
    with pytest.raises(Exception, match="Operating safe set is empty"):
        m.objective_argmax(torch.zeros(N))
    m2 = MarginalModel(0, domain, [GaussianDistribution(np.full(N, 5.0), np.eye(N))], beta=np.array([1.0]),
                       noise_rate=HomoscedasticNoise(1, [0.1]), use_objective_as_constraint=True)
    with pytest.raises(Exception, match="negative infinity"):
        m2.objective_argmax(torch.full((N,), -np.inf))
    with pytest.raises(Exception, match="NaN everywhere"):
        m2.objective_argmax(torch.full((N,), np.nan))
    x = m2.objective_argmax(torch.arange(N, dtype=torch.float64))
    assert float(x[0]) == 1.0


def test_step_with_general_m_and_points(cuda):
    """MarginalModel.step accepts m >= 0 observations given as points (SafeOpt /
    GoOSE replay)."""
    from lib.gp.kernels import stationary
    rng = np.random.default_rng(5)
    domain = rng.random((150, 2))
    mo, mc = _build_pair(domain, [O.Matern32(lengthscale=0.5)] * 2, [stationary.Matern32(lengthscale=0.5)] * 2,
                         beta=[2.0, 2.0], rho=[0.1, 0.2], X0=domain[[3]], y0=np.array([[0.2], [0.1]]),
                         fc_obj=False)
    idx = np.array([5, 9, 77])
    y = rng.standard_normal((2, 3))
    mo.step(domain[idx], y)
    mc.step(domain[idx], y)
    _check_state(mo, mc)
    mo.step(domain[[11]], y[:, :1]); mc.step(domain[[11]], y[:, :1])  # plain points, m = 1
    _check_state(mo, mc)
    mc.step(np.zeros((0, 2)), np.zeros((2, 0)))
    assert mc.t == mo.t == 4
    np.testing.assert_allclose(_np(mc.covariances), mo.covariances, rtol=1e-9, atol=1e-11)


def test_full_size_properties_n65536(cuda):
    """BASELINE config 3's N = 65,536 (fp32 storage to bound the test's memory
    and time): size-independent properties of the update and scoring kernels."""
    from talgp.engine import DeviceGP
    from talgp import _abi
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~20 GB of HBM")
    N = 65536
    g = np.arange(16) / 16.0
    domain = torch.as_tensor(np.stack(np.meshgrid(g, g, g, g, indexing="ij"), -1).reshape(-1, 4), device=cuda)
    gp = DeviceGP(N, 1, dtype=torch.float32, device=cuda)
    gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, 0.25, domain)
    gp.rho.fill_(0.1)
    gp.init_bounds([1.0])
    assert torch.all(gp.var == 1.0)
    rows = torch.tensor([0, 1234, 40000, 65535], device=cuda)
    before = gp.cov[0, rows, :N].clone()
    idx = 31337
    k = gp.cov[0, idx, :N].clone()
    gp.observe(idx, [0.3], [1.0])
    s = (k[idx].double() + 0.01).float()
    L = torch.sqrt(s)
    w = (k / L) / L
    want = before - k[rows][:, None] * w[None, :]
    assert torch.equal(gp.cov[0, rows, :N], want)  # exact closed form, sampled rows
    # symmetric up to rounding: k[a]*w[b] and k[b]*w[a] round differently, in the
    # reference's Kxt.T @ Sigma_inv_Kxt (gaussian_distribution.py:36) just the same
    torch.testing.assert_close(gp.cov[0, idx, :N], gp.cov[0, :, idx], rtol=1e-5, atol=1e-7)
    assert torch.equal(gp.var[0].float(), torch.diagonal(gp.cov[0, :, :N]))
    # scoring: VTL with a 4,096-point target set, rows variant == symmetric column variant
    A = torch.as_tensor(np.sort(np.random.default_rng(0).choice(N, 4096, replace=False)), device=cuda)
    F1 = gp.score_rows(_abi.RULE_VTL, A, 4096)
    F2 = gp.score_rows(_abi.RULE_VTL, A, 4096, sharded_cols=True)
    torch.testing.assert_close(F1, F2, rtol=1e-6, atol=0)  # fp32 storage: cov is symmetric to rounding only
    gp.masked_argmax(F1.clone(), None, 1, None)
    r = gp.read_result()
    assert r.status == 0 and float(F1[r.idx]) == float(F1.max()) == r.val
