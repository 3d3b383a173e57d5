"""GPU parity tests, kernel by kernel: the CUDA path (through the C ABI) against
the NumPy oracle on the same seeded inputs.

Tolerances (north_star): posterior means / variances within 1e-10 relative in
fp64 and 1e-4 in fp32; integer / index / mask work bit-exact."""
import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10
RTOL32 = 1e-4


def _dev(x, cuda, dtype=torch.float64):
    return torch.as_tensor(np.asarray(x), device=cuda).to(dtype).contiguous()


def _np(t):
    return t.detach().cpu().numpy()


def _problem(N=300, d=2, q=2, seed=0):
    rng = np.random.default_rng(seed)
    domain = rng.random((N, d))
    kernels = [O.Matern52(lengthscale=0.3), O.Gaussian(lengthscale=0.25), O.Matern32(lengthscale=0.4)][:q]
    covs = [k.covariance(domain) for k in kernels]
    return rng, domain, covs


def _engine(cuda, covs, rho, dtype=torch.float64, **kw):
    from talgp.engine import DeviceGP
    q, N = len(covs), covs[0].shape[0]
    gp = DeviceGP(N, q, dtype=dtype, device=cuda, **kw)
    for i, c in enumerate(covs):
        gp.load_covariance(i, torch.as_tensor(c))
    gp.rho.copy_(_dev(rho, cuda))
    return gp


def _oracle_model(domain, covs, rho, beta, fc_obj=False, cls=O.MarginalModel):
    q, N = len(covs), covs[0].shape[0]
    distrs = [O.GaussianDistribution(np.zeros(N), c.copy()) for c in covs]
    noise = O.HeteroscedasticNoise([(lambda X, i=i: rho[i][: X.shape[0]]) for i in range(q)])
    m = cls(0, domain, distrs, beta=np.asarray(beta, dtype=np.float64), noise_rate=noise,
            use_objective_as_constraint=fc_obj)
    m.noise_rate = type("N", (), {"at": staticmethod(lambda X: rho[:, : X.shape[0]]), "q": q})()
    return m


# --------------------------------------------------------------------------- gram
@pytest.mark.parametrize("kind", ["Gaussian", "Laplace", "Matern32", "Matern52", "Linear"])
@pytest.mark.parametrize("d", [1, 2, 4])
def test_gram_matches_oracle(cuda, kind, d):
    from lib.gp import kernels as K
    rng = np.random.default_rng(1)
    X = rng.random((133, d))
    Y = rng.random((57, d))
    if kind == "Linear":
        ko, kc = O.Linear(variance=1.7), K.non_stationary.Linear(variance=1.7)
    else:
        ko = getattr(O, kind)(variance=1.3, lengthscale=0.37)
        kc = getattr(K.stationary, kind)(variance=1.3, lengthscale=0.37)
    ref = ko.cross_covariance(X, Y)
    out = _np(kc.cross_covariance(X, Y))
    assert out.shape == (57, 133)  # [len(Y), len(X)] (base.py:20-22)
    np.testing.assert_allclose(out, ref, rtol=1e-13, atol=1e-15)
    out32 = _np(kc.cross_covariance(X, Y, dtype=torch.float32))
    np.testing.assert_allclose(out32, ref, rtol=RTOL32, atol=1e-6)


def test_gram_diagonal_exact_and_symmetric(cuda):
    from lib.gp.kernels import stationary
    X = np.linspace(-1, 1, 100).reshape(-1, 1)  # tests/test_discrete_model.py:13
    for cls in (stationary.Gaussian, stationary.Matern32, stationary.Matern52, stationary.Laplace):
        Kc = _np(cls(lengthscale=1).covariance(X))
        assert np.all(np.diag(Kc) == 1.0)
        assert np.array_equal(Kc, Kc.T)


def test_nearest_index_first_min(cuda):
    from lib.utils import get_indices, get_mask
    rng = np.random.default_rng(2)
    X = rng.random((500, 3))
    X[100] = X[7]  # duplicate: the first copy wins (quirk 6)
    A = np.concatenate([X[[7, 100, 3, 499]], rng.random((40, 3))])
    got = _np(get_indices(_dev(X, cuda), _dev(A, cuda)))
    assert np.array_equal(got, O.get_indices(X, A))
    assert np.array_equal(_np(get_mask(_dev(X, cuda), _dev(X, cuda))), O.get_mask(X, X))
    # view fast path
    Xd = _dev(X, cuda)
    assert _np(get_indices(Xd, Xd[42:43])).tolist() == [42]


# --------------------------------------------------------------------- rank-1 step
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("N", [300, 517])
def test_rank1_step_matches_oracle_fp64(cuda, variant, N):
    rng, domain, covs = _problem(N=N, q=2, seed=3)
    rho = np.stack([np.full(N, 0.1), 0.05 + 0.1 * rng.random(N)])  # heteroscedastic too
    beta = [2.0, 3.0]
    gp = _engine(cuda, covs, rho, rank1_variant=variant)
    gp.init_bounds(beta)
    model = _oracle_model(domain, covs, rho, beta)
    for t in range(6):
        idx = int(rng.integers(N))
        y = rng.standard_normal(2)
        gp.observe(idx, y, beta)
        model.step_idx(np.array([idx]), y[:, None], rho[:, [idx]])
        np.testing.assert_allclose(_np(gp.mean), model.means, rtol=RTOL64, atol=1e-14)
        np.testing.assert_allclose(_np(gp.var), model.variances, rtol=RTOL64, atol=1e-14)
        np.testing.assert_allclose(_np(gp.cov[:, :, :N]), model.covariances, rtol=RTOL64, atol=1e-14)
        np.testing.assert_allclose(_np(gp.l), model.l, rtol=RTOL64, atol=1e-13)
        np.testing.assert_allclose(_np(gp.u), model.u, rtol=RTOL64, atol=1e-13)
        # the incrementally kept variance IS the matrix diagonal, bit for bit
        assert torch.equal(gp.var, torch.diagonal(gp.cov[:, :, :N], dim1=1, dim2=2))
    assert torch.count_nonzero(gp.cov[:, :, N:]) == 0  # pad columns stay zero


def test_rank1_variants_bit_identical(cuda):
    rng, domain, covs = _problem(N=1000, q=1, seed=4)
    rho = np.full((1, 1000), 0.1)
    outs = []
    for variant in (1, 2, 3):
        gp = _engine(cuda, covs, rho, rank1_variant=variant)
        gp.init_bounds([1.0])
        r = np.random.default_rng(0)
        for t in range(4):
            gp.observe(int(r.integers(1000)), r.standard_normal(1), [1.0])
        outs.append(gp.cov.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_rank1_step_fp32(cuda):
    N = 400
    rng, domain, covs = _problem(N=N, q=2, seed=5)
    rho = np.full((2, N), 0.1)
    beta = [2.0, 2.0]
    gp = _engine(cuda, covs, rho, dtype=torch.float32)
    gp.init_bounds(beta)
    model = _oracle_model(domain, covs, rho, beta)
    for t in range(5):
        idx = int(rng.integers(N))
        y = rng.standard_normal(2)
        gp.observe(idx, y, beta)
        model.step_idx(np.array([idx]), y[:, None], rho[:, [idx]])
    np.testing.assert_allclose(_np(gp.mean), model.means, rtol=RTOL32, atol=1e-5)
    np.testing.assert_allclose(_np(gp.var), model.variances, rtol=RTOL32, atol=1e-5)
    np.testing.assert_allclose(_np(gp.cov[:, :, :N]).astype(np.float64), model.covariances,
                               rtol=RTOL32, atol=1e-5)
    assert torch.equal(gp.var.to(torch.float32), torch.diagonal(gp.cov[:, :, :N], dim1=1, dim2=2))


def test_observe_with_device_index_and_table(cuda):
    N = 200
    rng, domain, covs = _problem(N=N, q=1, seed=6)
    rho = np.full((1, N), 0.1)
    table = rng.standard_normal((1, N))
    a = _engine(cuda, covs, rho); a.init_bounds([1.0])
    b = _engine(cuda, covs, rho); b.init_bounds([1.0])
    a.observe(77, table[:, 77], [1.0])
    b.observe(torch.tensor([77], device=cuda), _dev(table, cuda), [1.0], y_gather=True, y_stride=N)
    assert torch.equal(a.cov, b.cov) and torch.equal(a.mean, b.mean) and torch.equal(a.l, b.l)


def test_negative_variance_gives_nan_bounds(cuda):
    """No clamping (quirk 4): a negative diagonal gives NaN stddev and NaN l/u."""
    N = 64
    cov = np.eye(N)
    cov[5, 5] = -1.0
    gp = _engine(cuda, [cov], np.full((1, N), 0.1)); gp.init_bounds([1.0])
    assert np.isnan(_np(gp.l)[0, 5]) and np.isnan(_np(gp.u)[0, 5])
    gp.observe(0, [0.5], [1.0])
    assert np.isnan(_np(gp.l)[0, 5]) and np.isnan(_np(gp.u)[0, 5])  # max/min propagate NaN


# ------------------------------------------------------------ m-observation update
@pytest.mark.parametrize("m", [1, 3, 10, 70])
def test_condition_matches_oracle_posterior(cuda, m):
    N = 350
    rng, domain, covs = _problem(N=N, q=1, seed=7)
    gp = _engine(cuda, covs, np.full((1, N), 0.1))
    obs_idx = rng.integers(N, size=m)
    if m >= 3:
        obs_idx[1] = obs_idx[0]  # sampling with replacement (prior.py:28)
    obs_idx = np.sort(obs_idx)
    y = rng.standard_normal(m)
    noise = np.full(m, 0.1)
    ref = O.GaussianDistribution(np.zeros(N), covs[0].copy()).posterior(obs_idx, y, noise)
    gp.condition(0, torch.as_tensor(obs_idx), torch.as_tensor(y), torch.as_tensor(noise))
    tol = dict(rtol=1e-9, atol=1e-11) if m <= 64 else dict(rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(_np(gp.mean[0]), ref.mean, **tol)
    np.testing.assert_allclose(_np(gp.cov[0, :, :N]), ref.covariance, **tol)
    assert torch.equal(gp.var[0], torch.diagonal(gp.cov[0, :, :N]))


def test_condition_empty_is_identity(cuda):
    N = 50
    rng, domain, covs = _problem(N=N, q=1, seed=8)
    gp = _engine(cuda, covs, np.full((1, N), 0.1))
    before = gp.cov.clone()
    gp.condition(0, torch.zeros((0,), dtype=torch.int64), torch.zeros((0,)), None)
    assert torch.equal(before, gp.cov)


def test_posterior_jitter_is_squared(cuda):
    """obs_noise_std=None adds jitter^2 = 1e-10, not 1e-5 (quirk 1)."""
    from lib.gp.gaussian_distribution import GaussianDistribution
    N = 40
    rng, domain, covs = _problem(N=N, q=1, seed=9)
    d = GaussianDistribution(np.zeros(N), covs[0])
    post = d.posterior(np.array([3, 9]), np.array([0.1, -0.2]))
    ref = O.GaussianDistribution(np.zeros(N), covs[0].copy()).posterior(np.array([3, 9]), np.array([0.1, -0.2]))
    np.testing.assert_allclose(_np(post.covariance), ref.covariance, rtol=0, atol=1e-9)
    np.testing.assert_allclose(_np(post.mean), ref.mean, rtol=1e-6, atol=1e-9)
    assert abs(_np(post.variance)[3]) < 1e-9  # conditioned (almost) exactly
    assert not torch.equal(d.covariance, post.covariance)  # out of place, like the reference


# ---------------------------------------------------------------------------- masks
@pytest.mark.parametrize("fc,q", [(0, 1), (1, 1), (1, 3), (0, 2)])
def test_bounds_masks_match_oracle(cuda, fc, q):
    N = 1000
    rng = np.random.default_rng(10 + fc + q)
    from talgp.engine import DeviceGP
    gp = DeviceGP(N, q, device=cuda)
    l = rng.standard_normal((q, N))
    u = l + rng.random((q, N))
    gp.l.copy_(_dev(l, cuda)); gp.u.copy_(_dev(u, cuda))
    pess, opt, pm, pe, max_l = gp.bounds_masks(fc)
    model = O.MarginalModel.__new__(O.MarginalModel)
    model._l, model._u, model.first_constraint, model.q = l, u, fc, q
    model.domain = np.zeros((N, 1))
    assert np.array_equal(_np(pess).astype(bool), model.pessimistic_safe_set)
    assert np.array_equal(_np(opt).astype(bool), model.optimistic_safe_set)
    assert _np(max_l)[0] == model.max_l
    assert np.array_equal(_np(pm).astype(bool), model.potential_maximizers)
    assert np.array_equal(_np(pe).astype(bool), model.potential_expanders)
    # empty safe set: max_l = -inf, every optimistic point is a potential maximizer
    if fc < q:
        gp.l.fill_(-1.0)
        pess, opt, pm, pe, max_l = gp.bounds_masks(fc)
        assert _np(max_l)[0] == -np.inf and int(pess.sum()) == 0
        assert torch.equal(pm, opt)
    # operating-safe-set override (UnsafeModel)
    ones = torch.ones((N,), dtype=torch.uint8, device=cuda)
    gp.l.copy_(_dev(l, cuda))
    _, _, _, _, max_l = gp.bounds_masks(fc, ones)
    assert _np(max_l)[0] == l[0].max()


@pytest.mark.parametrize("N", [1, 31, 4096, 4097, 100003])
def test_mask_to_indices(cuda, N):
    from talgp.engine import DeviceGP
    rng = np.random.default_rng(N)
    gp = DeviceGP(N, 1, device=cuda)
    for p in (0.0, 0.3, 1.0):
        mask = rng.random(N) < p
        idx, count = gp.mask_to_indices(torch.as_tensor(mask.astype(np.uint8), device=cuda))
        c = int(count.item())
        assert c == mask.sum()
        assert np.array_equal(_np(idx[:c]), np.nonzero(mask)[0])


def test_scalar_subset_matches_isin(cuda):
    from lib.model.marginal import ROI, MarginalModel
    from lib.model.unsafe import UnsafeModel  # noqa: F401
    from lib.gp.gaussian_distribution import GaussianDistribution
    from lib.noise import HomoscedasticNoise
    g = np.linspace(0, 1, 6)
    domain = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)  # 36 grid points
    N = len(domain)
    d = GaussianDistribution(np.zeros(N), np.eye(N))
    model = MarginalModel(0, domain, [d], beta=np.array([1.0]), noise_rate=HomoscedasticNoise(1, [0.1]),
                          use_objective_as_constraint=True)
    rng = np.random.default_rng(11)
    for trial in range(30):
        safe = rng.random(N) < rng.choice([0.0, 0.1, 0.5])
        roi = rng.random(N) < rng.choice([0.0, 0.15, 0.6])
        model._gp.l.copy_(_dev(np.where(safe, 1.0, -1.0)[None], cuda))
        r = ROI.from_mask(model, torch.as_tensor(roi, device=cuda))
        assert r.m == roi.sum()
        want = O.is_subset_of(domain[safe], domain[roi])
        assert model.safe_set_is_subset_of(r) == want
        assert model.safe_set_is_subset_of(_dev(domain[roi], cuda).reshape(-1, 2)) == want


# -------------------------------------------------------------------------- scoring
def _stepped_pair(cuda, N=400, q=2, steps=5, seed=12, dtype=torch.float64):
    rng, domain, covs = _problem(N=N, q=q, seed=seed)
    rho = np.stack([np.full(N, 0.1), 0.05 + 0.1 * rng.random(N), np.full(N, 0.2)][:q])
    beta = [2.0] * q
    gp = _engine(cuda, covs, rho, dtype=dtype)
    gp.init_bounds(beta)
    model = _oracle_model(domain, covs, rho, beta)
    for t in range(steps):
        idx = int(rng.integers(N))
        y = rng.standard_normal(q)
        gp.observe(idx, y, beta)
        model.step_idx(np.array([idx]), y[:, None], rho[:, [idx]])
    return rng, gp, model


def test_itl_undirected(cuda):
    rng, gp, model = _stepped_pair(cuda)
    np.testing.assert_allclose(_np(gp.score_itl_undirected()), O.itl_undirected(model), rtol=1e-12)


@pytest.mark.parametrize("m", [1, 7, 150, 400])
def test_vtl_rows_and_cols(cuda, m):
    from talgp import _abi
    rng, gp, model = _stepped_pair(cuda)
    roi = np.sort(rng.choice(gp.N, m, replace=False))
    ref = O.vtl_F(model, roi_idx=roi)
    ridx = torch.as_tensor(roi, device=cuda)
    F = _np(gp.score_rows(_abi.RULE_VTL, ridx, m))
    np.testing.assert_allclose(F, ref, rtol=RTOL64)
    Fc = _np(gp.score_rows(_abi.RULE_VTL, ridx, m, sharded_cols=True))  # symmetric column gather
    np.testing.assert_allclose(Fc, ref, rtol=RTOL64)


def test_vtl_fp32(cuda):
    from talgp import _abi
    rng, gp, model = _stepped_pair(cuda, dtype=torch.float32)
    roi = np.sort(rng.choice(gp.N, 60, replace=False))
    F = _np(gp.score_rows(_abi.RULE_VTL, torch.as_tensor(roi, device=cuda), 60))
    np.testing.assert_allclose(F, O.vtl_F(model, roi_idx=roi), rtol=RTOL32)


def test_ctl_and_mmitl(cuda):
    from talgp import _abi
    rng, gp, model = _stepped_pair(cuda, N=200)
    roi = np.sort(rng.choice(gp.N, 30, replace=False))
    A = model.domain[roi]
    sq = model.sqd_correlation(X=model.domain, A=A)
    ridx = torch.as_tensor(roi, device=cuda)
    np.testing.assert_allclose(_np(gp.score_rows(_abi.RULE_CTL, ridx, 30)), sq.sum(axis=(0, 1)), rtol=1e-10)
    with np.errstate(all="ignore"):
        ref = np.sum(-0.5 * np.log(1 - sq), axis=(0, 1))
    got = _np(gp.score_rows(_abi.RULE_MMITL, ridx, 30))
    fin = np.isfinite(ref)
    np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-8)


@pytest.mark.parametrize("N,m", [(300, 40), (700, 128), (1000, 300)])
def test_itl_directed_matches_oracle(cuda, N, m):
    rng, gp, model = _stepped_pair(cuda, N=N, q=2, steps=4, seed=N)
    roi = np.sort(rng.choice(N, m, replace=False))
    ref = O.itl_directed(model, roi_idx=roi)
    F = _np(gp.score_itl_directed(torch.as_tensor(roi, device=cuda), m))
    # the conditioning is ill-conditioned by construction (jitter^2 = 1e-10);
    # both sides are backward stable, the scores agree to ~1e-7
    np.testing.assert_allclose(F, ref, rtol=1e-6, atol=1e-8)
    assert np.argmax(F) == np.argmax(ref) or abs(ref[np.argmax(F)] - ref.max()) <= 1e-9 * max(1, abs(ref.max()))


@pytest.mark.parametrize("form,capacity", [("append", 512), ("append", 7), ("downdate", 0)])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_incremental_conditioning_matches_recompute(cuda, dtype, form, capacity):
    """Incremental target-set conditioning — the append form (csrc/condapp.cu; with a
    capacity of 7 observations it also re-bases every 7 steps) and the rank-1 downdate of
    L^-1 cov[A, :] (csrc/condinc.cu) — against re-conditioning from scratch at every step,
    over 60 observations, with the reference's jitter^2 = 1e-10 (ill-conditioned Sigma_AA)."""
    N, m, q = 1100, 200, 2
    rng = np.random.default_rng(21)
    domain = rng.random((N, 4))
    covs = [O.Matern52(lengthscale=0.25).covariance(domain), O.Gaussian(lengthscale=0.3).covariance(domain)]
    rho = np.stack([np.full(N, 0.1), 0.05 + 0.1 * rng.random(N)])
    beta = [1.0, 1.0]
    gp = _engine(cuda, covs, rho, dtype=dtype); gp.init_bounds(beta)
    gp.cond_form, gp.cond_capacity = form, capacity
    fresh = _engine(cuda, covs, rho, dtype=dtype); fresh.init_bounds(beta)
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
    key = object()
    tol = 1e-7 if dtype == torch.float64 else 2e-3
    for t in range(60):
        Fi = gp.score_itl_directed(roi, m, incremental=True, roi_key=key)
        Fr = fresh.score_itl_directed(roi, m)
        torch.testing.assert_close(Fi, Fr, rtol=tol, atol=tol * 1e-2)
        idx = int(torch.argmax(Fr)) if t % 2 else int(rng.integers(N))
        y = rng.standard_normal(q)
        gp.observe(idx, y, beta)
        fresh.observe(idx, y, beta)
    assert torch.equal(gp.cov, fresh.cov)
    # a general-m update invalidates the resident factor; the next score re-conditions
    for g in (gp, fresh):
        g.condition(0, torch.tensor([5, 9]), torch.tensor([0.1, 0.2]), torch.tensor([0.1, 0.1]))
    torch.testing.assert_close(gp.score_itl_directed(roi, m, incremental=True, roi_key=key),
                               fresh.score_itl_directed(roi, m), rtol=tol, atol=tol * 1e-2)


@pytest.mark.parametrize("chunks", [2, 8, 32])
def test_cond_update_row_parallel_matches_single_pass(cuda, chunks):
    """The prefix-sum (row-parallel) form of the downdate against the one-pass recurrence."""
    import ctypes as C
    from talgp import _abi
    lib = _abi.load()
    rng = np.random.default_rng(5)
    m, n = 512, 300
    Vh = rng.standard_normal((m, n)) * 0.03  # ||p||^2 = ||V[:, x]||^2 / s ~ 0.36 < 1
    Vh[400:] = 0.0  # pad rows
    w = rng.standard_normal(n) * 0.3
    k = np.ones(n); rho = np.full(n, 0.1)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    kd, wd, rd = _dev(k, cuda), _dev(w, cuda), _dev(rho, cuda)  # keep the device buffers alive
    for ch in (1, chunks):
        V = _dev(Vh, cuda)
        pcol = V[:, 17].clone()
        coef = torch.empty((int(lib.tal_cond_coef_bytes(m)) // 8,), dtype=torch.float64, device=cuda)
        scratch = torch.empty((int(lib.tal_cond_chunk_scratch_bytes(n)) // 8,), dtype=torch.float64, device=cuda)
        cs = torch.empty((n,), dtype=torch.float64, device=cuda)
        _abi.check(lib.tal_cond_update(p(V), n, m, n, n - 4, 0, p(pcol), p(kd), p(wd),
                                       p(rd), None, 17, p(coef), p(cs), ch, p(scratch), _abi.F64, st))
        outs.append((V, cs))
    assert torch.isfinite(outs[0][0]).all() and torch.isfinite(outs[1][0]).all()
    torch.testing.assert_close(outs[1][0], outs[0][0], rtol=1e-11, atol=1e-14)
    torch.testing.assert_close(outs[1][1], outs[0][1], rtol=1e-11, atol=1e-14)


def test_potrf_trsm_against_lapack(cuda):
    """The DMMA Cholesky / TRSM on a well-conditioned SPD matrix: 1e-11."""
    import ctypes as C
    import scipy.linalg as sl
    from talgp import _abi
    lib = _abi.load()
    rng = np.random.default_rng(13)
    m, n = 384, 256
    A = rng.standard_normal((m, m))
    S = A @ A.T + m * np.eye(m)
    B = rng.standard_normal((m, n))
    Sd, Bd = _dev(S, cuda), _dev(B, cuda)
    work = torch.empty((int(lib.tal_potrf_work_bytes(m)) // 8,), dtype=torch.float64, device=cuda)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    _abi.check(lib.tal_potrf_lower(p(Sd), m, m, p(work), st))
    _abi.check(lib.tal_trsm_lower(p(Sd), m, m, p(work), p(Bd), n, n, st))
    L = np.linalg.cholesky(S)
    np.testing.assert_allclose(np.tril(_np(Sd)), L, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(_np(Bd), sl.solve_triangular(L, B, lower=True), rtol=1e-10, atol=1e-12)
    # not positive definite -> NaN, like jnp.linalg.cholesky
    bad = _dev(-np.eye(128), cuda)
    _abi.check(lib.tal_potrf_lower(p(bad), 128, 128, p(work), st))
    assert torch.isnan(bad[0, 0])


# --------------------------------------------------------------------------- argmax
def test_masked_argmax_matches_oracle_tie_set(cuda):
    from talgp.engine import DeviceGP
    from talgp import _abi
    N = 5000
    rng = np.random.default_rng(14)
    gp = DeviceGP(N, 2, device=cuda)
    l = rng.standard_normal((2, N))
    gp.l.copy_(_dev(l, cuda))
    model = O.MarginalModel.__new__(O.MarginalModel)
    model._l, model.first_constraint, model.q = l, 1, 2
    F = np.round(rng.standard_normal(N), 1)  # many exact ties
    F[rng.integers(N, size=50)] = np.nan
    sample = rng.random(N) < 0.7
    Fm = F.copy(); Fm[~sample] = -np.inf
    ties, mx = model.objective_tie_set(Fm)
    Fd = _dev(F, cuda)
    gp.masked_argmax(Fd, None, 1, torch.as_tensor(sample.astype(np.uint8), device=cuda))
    res = gp.read_result()
    assert res.status == _abi.ARGMAX_OK
    assert res.idx == ties[0] and res.n_ties == len(ties) and res.val == mx
    np.testing.assert_array_equal(_np(Fd), Fm)  # F was masked by the sample region in place


def test_masked_argmax_statuses(cuda):
    from talgp.engine import DeviceGP
    from talgp import _abi
    N = 300
    gp = DeviceGP(N, 1, device=cuda)
    F = torch.zeros((N,), dtype=torch.float64, device=cuda)
    gp.l.fill_(-1.0)
    gp.masked_argmax(F.clone(), None, 0, None)
    assert gp.read_result().status == _abi.ARGMAX_EMPTY_SAFE_SET
    gp.l.fill_(1.0)
    gp.masked_argmax(torch.full_like(F, -np.inf), None, 0, None)
    assert gp.read_result().status == _abi.ARGMAX_ALL_NEG_INF
    gp.l.fill_(-1.0); gp.l[0, 10:20] = 1.0
    Fn = F.clone(); Fn[10:20] = float("nan")
    gp.masked_argmax(Fn, None, 0, None)
    assert gp.read_result().status == _abi.ARGMAX_ALL_NAN
    # all safe entries -inf but F not all -inf: ties include the unsafe points (quirk 7)
    Fi = F.clone(); Fi[10:20] = -np.inf
    gp.masked_argmax(Fi, None, 0, None)
    r = gp.read_result()
    assert r.status == _abi.ARGMAX_OK and r.val == -np.inf and r.idx == 0 and r.n_ties == N
    # fc == q: no constraints, everything is safe
    gp.masked_argmax(torch.arange(N, dtype=torch.float64, device=cuda), None, 1, None)
    assert gp.read_result().idx == N - 1


def test_argmax_combine(cuda):
    import ctypes as C
    from talgp import _abi
    lib = _abi.load()
    parts = (_abi.ArgmaxResult * 3)()
    parts[0] = _abi.ArgmaxResult(40, 1.5, 2, 0, 7)
    parts[1] = _abi.ArgmaxResult(12, 1.5, 1, 0, 7)
    parts[2] = _abi.ArgmaxResult(-1, -np.inf, 0, 2, 0)
    buf = torch.frombuffer(bytearray(bytes(parts)), dtype=torch.uint8).to(cuda)
    out = torch.zeros((32,), dtype=torch.uint8, device=cuda)
    _abi.check(lib.tal_argmax_combine(C.c_void_p(buf.data_ptr()), 3, C.c_void_p(out.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    r = _abi.ArgmaxResult.from_buffer_copy(out.cpu().numpy().tobytes())
    assert (r.idx, r.val, r.n_ties, r.status) == (12, 1.5, 3, 0)
