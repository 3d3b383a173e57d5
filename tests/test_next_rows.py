"""SURVEY 8f-2 (ContinuousModel / LineBO) and 8f-4 (Thompson-sampling ROI) against golden
vectors produced by the reference's own source (tests/golden/reference_next.npz, made by
tests/golden/make_reference_fixtures.py over the NumPy stand-in for jax): the oracle on the
CPU tier, the CUDA path through the drop-in `lib.*` classes on the GPU tier."""
import os

import numpy as np
import pytest

NEXT = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_next.npz"))
TIGHT = dict(rtol=1e-10, atol=1e-12)


class OracleBE:
    def __init__(self):
        import oracle as O
        self.O = O
        self.ContinuousModel, self.LineBO, self.UnknownFunction = O.ContinuousModel, O.LineBO, O.UnknownFunction
        self.HomoscedasticNoise, self.ITL, self.pm = O.HomoscedasticNoise, O.ITL, O.pm_roi_constructor
        self.xp = np

    def kernel(self, name, **kw):
        return getattr(self.O, name)(**kw)

    def arr(self, x):
        return np.asarray(x, dtype=np.float64)

    def np(self, x):
        return np.asarray(x)

    def directed(self, Alg, roi):
        return lambda model: Alg(model, roi)

    def thompson(self, samples, I, J):
        J = np.zeros(samples.shape[1], dtype=bool) if J is None else np.asarray(J)
        return self.O.thompson_masks_from_samples(samples, np.asarray(I), J)


class CudaBE:
    def __init__(self):
        import torch
        from lib.algorithms import pm_roi_constructor
        from lib.algorithms.itl import ITL
        from lib.algorithms.line_bo import LineBO, directed_alg_constructor
        from lib.function.unknown import UnknownFunction
        from lib.gp.kernels import stationary
        from lib.model.continuous import ContinuousModel
        from lib.noise import HomoscedasticNoise
        self.torch, self.stationary = torch, stationary
        self.ContinuousModel, self.LineBO, self.UnknownFunction = ContinuousModel, LineBO, UnknownFunction
        self.HomoscedasticNoise, self.ITL, self.pm = HomoscedasticNoise, ITL, pm_roi_constructor
        self.directed = directed_alg_constructor
        self.xp = torch

    def kernel(self, name, **kw):
        return getattr(self.stationary, name)(**kw)

    def arr(self, x):
        return self.torch.as_tensor(np.asarray(x, dtype=np.float64), device="cuda")

    def np(self, x):
        return x.detach().cpu().numpy() if isinstance(x, self.torch.Tensor) else np.asarray(x)

    def thompson(self, samples, I, J):
        from lib.model.marginal import MarginalModel
        t = self.torch
        J = t.zeros(samples.shape[1], dtype=t.bool) if J is None else t.as_tensor(np.asarray(J))
        m, v = MarginalModel.thompson_masks_from_samples(self.arr(samples), t.as_tensor(np.asarray(I)), J)
        return self.np(m), self.np(v)


@pytest.fixture(scope="module")
def oracle_be():
    return OracleBE()


@pytest.fixture(scope="module")
def cuda_be(cuda):
    return CudaBE()


@pytest.fixture(params=[pytest.param("oracle_be", id="oracle"),
                        pytest.param("cuda_be", id="cuda", marks=pytest.mark.gpu)])
def be(request):
    return request.getfixturevalue(request.param)


def test_continuous_model_marginalize_matches_reference(be):
    """lib/model/continuous.py:53-95 — before any observation and after 7 (one location repeated)."""
    noise = be.HomoscedasticNoise(2, be.arr([0.1, 0.2]))
    ks = [be.kernel("Gaussian", lengthscale=0.8), be.kernel("Matern52", variance=1.5, lengthscale=0.6)]
    cm = be.ContinuousModel(0, d=2, prior_kernels=ks, beta=np.array([2.0, 3.0]), noise_rate=noise)
    Xq = be.arr(NEXT["cont/Xq"])
    m0 = cm.marginalize(Xq)
    for attr in ("means", "covariances", "l", "u"):
        np.testing.assert_allclose(be.np(getattr(m0, attr)), NEXT[f"cont/t0/{attr}"], **TIGHT)
    cm.step(be.arr(NEXT["cont/Xa"]), be.arr(NEXT["cont/ya"]))
    cm.step(be.arr(NEXT["cont/Xb"]), be.arr(NEXT["cont/yb"]))
    m7 = cm.marginalize(Xq)
    assert int(m7.t) == int(NEXT["cont/t7/t"]) == 7
    for attr in ("means", "covariances", "l", "u"):
        np.testing.assert_allclose(be.np(getattr(m7, attr)), NEXT[f"cont/t7/{attr}"], rtol=1e-9, atol=1e-11)
    assert np.array_equal(be.np(m7.pessimistic_safe_set).astype(bool), NEXT["cont/t7/pess"])


def test_line_bo_matches_reference(be):
    """Five LineBO steps along given directions with a nested ITL-PM (lib/algorithms/line_bo.py:
    80-163): sub-space bounds, the extended domain, the nested model, the chosen point."""
    dirs, bounds, x0 = NEXT["linebo/dirs"], NEXT["linebo/bounds"], NEXT["linebo/x0"]
    xp, c = be.xp, be.arr([0.4, 0.6])
    f = be.UnknownFunction(q=1, f=lambda x: xp.exp(-xp.sum((x - c) ** 2)).reshape(1),
                           use_objective_as_constraint=True)
    noise = be.HomoscedasticNoise(1, be.arr([0.05]))
    lm = be.ContinuousModel(1, d=2, prior_kernels=[be.kernel("Gaussian", lengthscale=0.5)], beta=np.array([2.0]),
                            noise_rate=noise, use_objective_as_constraint=True)
    lm.step(be.arr(x0), f.observe(be.arr(x0)))
    state = {"it": 0}

    class FixedLineBO(be.LineBO):
        def _direction(self, f):
            v = be.arr(dirs[state["it"]])
            state["it"] += 1
            return v

    lbo = FixedLineBO(alg=be.directed(be.ITL, be.pm), model=lm, bounds=be.arr(bounds), n=12, eps=0.01)
    for t in range(len(dirs)):
        res = lbo.step(f)
        mm = res["marginal_model"]
        np.testing.assert_allclose(be.np(mm.domain), NEXT[f"linebo/{t}/domain"], rtol=1e-12, atol=1e-14)
        sb = [float(b) for b in lbo.subspace_bounds()]
        np.testing.assert_allclose(sb, NEXT[f"linebo/{t}/subspace_bounds"], rtol=1e-12)
        np.testing.assert_allclose(be.np(mm.means), NEXT[f"linebo/{t}/means"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(be.np(mm.variances), NEXT[f"linebo/{t}/variances"], rtol=1e-8, atol=1e-10)
        Fn, Fg = be.np(res["F"]), NEXT[f"linebo/{t}/F"]
        fin = np.isfinite(Fg)
        np.testing.assert_allclose(Fn[fin], Fg[fin], rtol=1e-5, atol=1e-8)
        # the extended domain holds duplicates (observed points, earlier optima): ties between copies of
        # the SAME point are broken differently (reference: jr.choice), the chosen POINT is the same
        np.testing.assert_allclose(be.np(res["x"]), NEXT[f"linebo/{t}/x"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(be.np(res["y"]), NEXT[f"linebo/{t}/y"], rtol=1e-12)
        np.testing.assert_allclose(be.np(res["argmax"]), NEXT[f"linebo/{t}/argmax"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(be.np(lm.X), NEXT["linebo/final_X"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(be.np(lm.y), NEXT["linebo/final_y"], rtol=1e-12)
    np.testing.assert_allclose(be.np(lbo.encountered_max), NEXT["linebo/encountered_max"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("name,I,J", [("obj0_cons12", [0], [1, 2]), ("obj01_cons2", [0, 1], [2]),
                                      ("obj0_nocons", [0], None)])
def test_thompson_masks_match_reference(be, name, I, J):
    """marginal.py:238-312 on preset samples (one of them without any safe point)."""
    masks, values = be.thompson(NEXT["ts/samples"], I, J)
    assert np.array_equal(masks.any(axis=0), NEXT[f"ts/{name}/mask"])
    np.testing.assert_array_equal(values, NEXT[f"ts/{name}/values"])


@pytest.mark.gpu
def test_sampling_statistics_and_ts_roi(cuda):
    """GaussianDistribution.sample / MarginalModel.sample: the empirical moments of 20,000 draws
    match mean and covariance; the Thompson-sampling ROI constructor returns a non-empty subset
    of the domain and ITL-TS takes a step."""
    import torch
    from lib.algorithms import ts_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.function import TableFunction
    from lib.gp.gaussian_distribution import GaussianDistribution
    from lib.gp.kernels import stationary
    from lib.gp.means import ZeroMean
    from lib.gp.prior import prior_distr
    from lib.model.marginal import MarginalModel
    from lib.noise import HomoscedasticNoise
    rng = np.random.default_rng(0)
    N = 120
    dom = np.sort(rng.random((N, 1)) * 4, axis=0)
    noise = HomoscedasticNoise(2, np.array([0.1, 0.1]))
    table = np.stack([np.sin(2 * dom[:, 0]), 0.8 - 0.1 * dom[:, 0]]) + 0.05 * rng.standard_normal((2, N))
    seed = np.array([5, 40, 90])
    distrs = prior_distr(noise, dom, dom[seed], table[:, seed],
                         [stationary.Gaussian(lengthscale=0.7), stationary.Matern52(lengthscale=1.0)],
                         [ZeroMean()] * 2)
    model = MarginalModel(3, dom, distrs, beta=np.array([2.0, 2.0]), noise_rate=noise)
    d0 = model.distr(0)
    s = d0.sample(key=11, sample_shape=(20000,))
    assert s.shape == (20000, N)
    torch.testing.assert_close(s.mean(dim=0), d0.mean, rtol=0, atol=0.03)
    emp = torch.cov(s.T)
    torch.testing.assert_close(emp, d0.covariance, rtol=0, atol=0.04)
    assert model.sample(k=7).shape == (2, 7, N)
    mask = model.thompson_sampling(k=16, J=torch.arange(1, 2))
    assert mask.dtype == torch.bool and 0 < int(mask.sum()) <= 16
    vals = model.thompson_sampling_values(k=16, J=torch.arange(1, 2))
    assert vals.shape == (16,) and bool(torch.isfinite(vals).any())
    alg = ITL(model, ts_roi_constructor(8))
    res = alg.step(TableFunction(model.domain, table))
    assert res["F"].shape == (N,) and model.t == 1


# ---- the reference's own tests, restated over the drop-in classes ---------------------------------
@pytest.mark.gpu
def test_reference_test_continuous_model(cuda):
    """tests/test_continuous_model.py of the reference: prior state of a marginalised model is exact,
    sample() has the documented shape, step() changes the marginal means."""
    import torch
    from lib.gp import kernels
    from lib.model.continuous import ContinuousModel
    from lib.noise import HomoscedasticNoise
    noise_rate = HomoscedasticNoise(q=1, noise_rates=np.array([0.1]))
    model = ContinuousModel(key=0, d=1, prior_kernels=[kernels.stationary.Gaussian(lengthscale=1)],
                            beta=np.array([1.0]), noise_rate=noise_rate, use_objective_as_constraint=True)
    X = torch.tensor([[-1.0], [-0.5], [0.0], [0.5], [1.0]], dtype=torch.float64, device=cuda)
    N = X.shape[0]
    mm = model.marginalize(X)
    assert bool((mm.means == torch.zeros((1, N), device=cuda, dtype=torch.float64)).all())      # :30-31
    assert bool((mm.stddevs == torch.ones((1, N), device=cuda, dtype=torch.float64)).all())     # :34-35
    assert bool((mm.l == -1.0).all()) and bool((mm.u == 1.0).all())                             # :38-40
    assert mm.sample(k=10).shape == (1, 10, N)                                                  # :43-44
    model.step(X, torch.sin(X).T)                                                               # :47-51
    assert bool((model.marginalize(X).means != 0).any())


@pytest.mark.gpu
def test_reference_test_line_bo(cuda):
    """tests/test_line_bo.py of the reference: Random / Coordinate / Descent LineBO return unit-length
    directions."""
    import torch
    from lib.algorithms import DiscreteGreedyAlgorithm
    from lib.algorithms.line_bo import CoordinateLineBO, DescentLineBO, RandomLineBO, alg_constructor
    from lib.function.unknown import UnknownFunction
    from lib.gp import kernels
    from lib.model.continuous import ContinuousModel
    from lib.noise import HomoscedasticNoise

    class DummyAlg(DiscreteGreedyAlgorithm):
        def F(self):
            pass

    D = 2
    model = ContinuousModel(key=0, d=D, prior_kernels=[kernels.stationary.Gaussian()], beta=np.array([1.0]),
                            noise_rate=HomoscedasticNoise(q=1, noise_rates=np.array([0.1])),
                            use_objective_as_constraint=True)
    bounds = np.array([[-1.0, 1.0], [0.0, 1.0]])
    f = UnknownFunction(q=1, f=lambda x: torch.linalg.norm(x).reshape(1))
    kw = dict(alg=alg_constructor(DummyAlg), model=model, bounds=bounds, n=10, eps=0.01)
    for lbo in (RandomLineBO(**kw), CoordinateLineBO(**kw), DescentLineBO(m=2 * D, alpha=0.1, **kw)):
        assert float(torch.linalg.norm(lbo._normalized_direction(f))) == pytest.approx(1, abs=1e-5)


def test_truvar_matches_reference(be):
    """TruVar (lib/algorithms/tru_var/__init__.py:17-50) with TruVarModel's shrinking
    variance threshold (model.py:28-33) on the fixed-ROI problem c3: F, eta and the selections over
    8 steps, teacher-forced."""
    import problems
    p = problems.c3(T=8)
    A = np.asarray(p["roi"])
    if isinstance(be, OracleBE):
        O = be.O
        nz = O.HomoscedasticNoise(1, np.asarray(p["rho"]))
        distrs = O.prior_distr(nz, p["domain"], p["X0"], p["y0"], [O.Matern52(variance=1.0, lengthscale=0.25)],
                               [O.ZeroMean()])
        model = O.TruVarModel(eta=1.2, delta=0.3, r=0.5, key=4, domain=p["domain"], distrs=distrs,
                              beta=np.array([2.0]), noise_rate=nz)
        alg = O.TruVar(model, lambda m: m.domain[A])
        f = O.TableFunction(p["domain"], p["table"])
    else:
        from lib.algorithms import index_roi_constructor
        from lib.algorithms.tru_var import TruVar
        from lib.algorithms.tru_var.model import TruVarModel
        from lib.function import TableFunction
        from lib.gp.means import ZeroMean
        from lib.gp.prior import prior_distr
        nz = be.HomoscedasticNoise(1, np.asarray(p["rho"]))
        distrs = prior_distr(nz, p["domain"], p["X0"], p["y0"], [be.kernel("Matern52", variance=1.0, lengthscale=0.25)],
                             [ZeroMean()])
        model = TruVarModel(eta=1.2, delta=0.3, r=0.5, key=4, domain=p["domain"], distrs=distrs,
                            beta=np.array([2.0]), noise_rate=nz)
        alg = TruVar(model, index_roi_constructor(A))
        f = TableFunction(model.domain, p["table"])
    for t in range(8):
        F = be.np(alg.F())
        assert float(model.eta) == pytest.approx(float(NEXT["truvar/eta"][t]), rel=1e-12)
        np.testing.assert_allclose(F, NEXT["truvar/F"][t], rtol=1e-10, atol=1e-12)
        idx = int(NEXT["truvar/sel"][t])
        assert F[idx] >= F.max() - 1e-9 * max(1.0, abs(F.max()))
        X = model.domain[idx: idx + 1]
        model.step(X, f.observe(X))
    np.testing.assert_allclose(be.np(model.means), NEXT["truvar/means"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(be.np(model.variances), NEXT["truvar/variances"], rtol=1e-10, atol=1e-12)
