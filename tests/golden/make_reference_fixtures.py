"""Golden vectors produced by the REFERENCE'S OWN SOURCE (/root/reference/lib),
imported and executed in the build container over the NumPy stand-in for jax in
tests/golden/refshim.py (jax/jaxlib 0.4.30 are not installable here):

    python tests/golden/make_reference_fixtures.py

writes
  tests/golden/reference_trajectories.npz   BO trajectories of the problems in
      tests/problems.py through lib.gp.prior.prior_distr, lib.model.marginal.
      MarginalModel, lib.algorithms.{itl.ITL, vtl.VTL, ctl.CTL, mm_itl.MMITL} and
      the ROI constructors: per step the masked F (lib/algorithms/__init__.py:80),
      the tie set's lowest index and size, the index the reference's
      objective_argmax chose, the safe sets / PM / PE masks; final means,
      variances, l, u.  Teacher-forced on the lowest tied index (exact ties are
      broken by jr.choice in the reference, marginal.py:346-347).
  tests/golden/reference_calls.npz          inputs and outputs of single calls:
      kernels, get_indices / get_mask / is_subset_of, compute_posterior_covariance,
      GaussianDistribution.posterior (jitter and noise paths, repeated indices),
      MarginalModel.step with m = 0, 1, 5, sqd_correlation, the error messages of
      objective_argmax.
  tests/golden/reference_metrics.npz        SURVEY 8f-4 remainder: entropy / masked_entropy, information_gain,
      the information ratios and the per-step transductive-learning metrics on a 2-GP model after 3 + 4
      observations.
  tests/golden/reference_next.npz           SURVEY 8f-2 / 8f-4: ContinuousModel.marginalize before
      and after 7 observations, five LineBO steps along given directions (nested ITL-PM),
      thompson_sampling / thompson_sampling_values on preset samples.

This script needs /root/reference and is NOT run on the GPU box; the .npz files
are committed.  The first thing it does is run the reference's own tests over the
stand-in (tests/test_discrete_model.py, test_continuous_model.py, test_line_bo.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("TAL_REFERENCE", "/root/reference")

import refshim  # noqa: E402

refshim.install()
sys.path.insert(0, REF)

import jax.numpy as jnp  # noqa: E402  (the stand-in)
import jax.random as jr  # noqa: E402
from lib.algorithms import (  # noqa: E402
    f_roi_constructor, fixed_roi_constructor, pe_roi_constructor, pm_roi_constructor, ucb_roi_constructor)
from lib.algorithms.ctl import CTL  # noqa: E402
from lib.algorithms.itl import ITL  # noqa: E402
from lib.algorithms.mm_itl import MMITL  # noqa: E402
from lib.algorithms.vtl import VTL  # noqa: E402
from lib.function import Function  # noqa: E402
from lib.gp.gaussian_distribution import GaussianDistribution, compute_posterior_covariance  # noqa: E402
from lib.gp.kernels import stationary  # noqa: E402
from lib.gp.means import ZeroMean  # noqa: E402
from lib.gp.prior import prior_distr  # noqa: E402
from lib.model.marginal import MarginalModel  # noqa: E402
from lib.model.unsafe import UnsafeModel  # noqa: E402
from lib.noise import HeteroscedasticNoise, HomoscedasticNoise  # noqa: E402
from lib.utils import get_indices, get_mask, is_subset_of  # noqa: E402

import problems  # noqa: E402


class TableFunction(Function):
    """Black box backed by a table of pre-drawn noisy observations (host)."""

    def __init__(self, domain, table, first_constraint):
        self.domain, self.table = domain, table
        self.q, self.first_constraint = table.shape[0], first_constraint

    def observe(self, X):
        idx = np.asarray(get_indices(self.domain, X))
        return jnp.array(self.table[:, idx])


RULES = {"ITL": ITL, "VTL": VTL, "CTL": CTL, "MMITL": MMITL}


def build_reference(p, rule=None):
    q = len(p["kernels"])
    domain = jnp.array(p["domain"])
    ks = [getattr(stationary, k)(variance=v, lengthscale=l) for k, v, l in p["kernels"]]
    noise = HomoscedasticNoise(q, jnp.array(p["rho"]))
    distrs = prior_distr(noise, domain, jnp.array(p["X0"]), jnp.array(p["y0"]), ks, [ZeroMean()] * q)
    model = MarginalModel(jr.PRNGKey(0), domain, distrs, beta=jnp.array(np.asarray(p["beta"], float)),
                          noise_rate=noise, use_objective_as_constraint=p["fc_obj"])
    if isinstance(p["roi"], str):
        ctor = {"pm": pm_roi_constructor, "pe": pe_roi_constructor, "f": f_roi_constructor,
                "ucb": ucb_roi_constructor}[p["roi"]]
    else:
        A = np.asarray(p["roi"])
        ctor = lambda m: m.domain[A]  # noqa: E731  a fixed target set, as points
    alg = RULES[rule or p["rule"]](model, ctor)
    return model, alg, TableFunction(domain, p["table"], model.first_constraint)


def run_trajectory(p, rule=None):
    model, alg, f = build_reference(p, rule)
    rec = {k: [] for k in ("F", "sel", "n_ties", "ref_choice", "pess", "opt", "pm", "pe", "roi_size", "max_l")}
    for t in range(p["T"]):
        rec["pess"].append(np.asarray(model.pessimistic_safe_set))
        rec["opt"].append(np.asarray(model.optimistic_safe_set))
        rec["pm"].append(np.asarray(model.potential_maximizers))
        rec["pe"].append(np.asarray(model.potential_expanders))
        rec["max_l"].append(float(model.max_l))
        rec["roi_size"].append(int(alg.roi.shape[0]))
        # DiscreteGreedyAlgorithm._step, lib/algorithms/__init__.py:80-81
        F = alg.F().at[~get_mask(model.domain, alg.sample_region)].set(-jnp.inf)
        Xc = model.objective_argmax(F=F).reshape(1, -1)
        choice = int(model.get_indices(Xc)[0])
        safe_F = np.where(np.asarray(model.operating_safe_set), np.asarray(F), -np.inf)
        ties = np.flatnonzero(safe_F == np.nanmax(safe_F))
        assert choice in ties
        idx = int(ties[0])
        rec["F"].append(np.asarray(F)); rec["sel"].append(idx); rec["n_ties"].append(len(ties))
        rec["ref_choice"].append(choice)
        X = model.domain[idx: idx + 1]
        model.step(X=X, y=f.observe(X))
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(means=np.asarray(model.means), variances=np.asarray(model.variances), l=np.asarray(model.l),
               u=np.asarray(model.u), t=np.array(model.t))
    return out


def _spd(rng, n):
    X = rng.random((n, 3))
    return np.asarray(stationary.Matern32(lengthscale=0.7).covariance(jnp.array(X))), X


def single_calls():
    out = {}
    rng = np.random.default_rng(42)
    # kernels: cross_covariance(X, Y) is [len(Y), len(X)] (lib/gp/kernels/base.py:17-25)
    X, Y = rng.random((9, 3)) * 2 - 1, rng.random((5, 3)) * 2 - 1
    out["kernel/X"], out["kernel/Y"] = X, Y
    for name in ("Gaussian", "Laplace", "Matern32", "Matern52"):
        k = getattr(stationary, name)(variance=1.7, lengthscale=0.6)
        out[f"kernel/{name}/cross"] = np.asarray(k.cross_covariance(jnp.array(X), jnp.array(Y)))
        out[f"kernel/{name}/cov"] = np.asarray(k.covariance(jnp.array(X)))
    # get_indices / get_mask / is_subset_of (lib/utils.py:53-71), with duplicate points and a tie
    D = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [0.5, 0.5], [2.0, 2.0]])
    A = np.array([[1.0, 0.0], [0.5, 0.5], [0.5, 0.0], [2.0, 2.1]])
    out["index/D"], out["index/A"] = D, A
    out["index/get_indices"] = np.asarray(get_indices(jnp.array(D), jnp.array(A)))
    out["index/get_mask"] = np.asarray(get_mask(jnp.array(D), jnp.array(A)))
    S1 = np.array([[0.0, 1.0], [1.0, 0.0]])
    S2 = np.array([[0.0, 0.5], [3.0, 0.0]])
    out["subset/S1"], out["subset/S2"] = S1, S2
    out["subset/S1_in_D"] = np.array(bool(is_subset_of(jnp.array(S1), jnp.array(D))))
    out["subset/S2_in_D"] = np.array(bool(is_subset_of(jnp.array(S2), jnp.array(D))))
    out["subset/rowwise_false_scalarwise_true"] = np.array(
        bool(is_subset_of(jnp.array([[1.0, 1.0]]), jnp.array(D))))  # quirk: scalar-wise isin
    out["subset/empty_in_D"] = np.array(bool(is_subset_of(jnp.zeros((0, 2)), jnp.array(D))))
    out["subset/D_in_empty"] = np.array(bool(is_subset_of(jnp.array(D), jnp.zeros((0, 2)))))
    # compute_posterior_covariance / posterior (lib/gp/gaussian_distribution.py:16-121)
    K, Xk = _spd(rng, 24)
    mu = rng.standard_normal(24)
    out["post/K"], out["post/mu"] = K, mu
    cases = {
        "jitter5": (np.array([3, 11, 17, 20, 5]), None),
        "noise5": (np.array([3, 11, 17, 20, 5]), np.array([0.1, 0.2, 0.05, 0.3, 0.1])),
        "repeat": (np.array([4, 4, 9]), np.array([0.1, 0.1, 0.2])),
        "scalar_noise": (np.array([7, 2]), 0.25),
        "single": (np.array([13]), np.array([0.1])),
        "empty": (np.zeros((0,), dtype=np.int64), None),
    }
    for name, (oi, ns) in cases.items():
        y = rng.standard_normal(len(oi))
        nsj = None if ns is None else (ns if np.isscalar(ns) else jnp.array(ns))
        cov = compute_posterior_covariance(jnp.array(K), jnp.array(oi), nsj)
        d = GaussianDistribution(mean=jnp.array(mu), covariance=jnp.array(K)).posterior(
            obs_idx=jnp.array(oi), obs=jnp.array(y), obs_noise_std=nsj)
        out[f"post/{name}/obs_idx"], out[f"post/{name}/y"] = oi, y
        out[f"post/{name}/noise"] = np.array(np.nan) if ns is None else np.asarray(ns, dtype=float)
        out[f"post/{name}/cov_only"] = np.asarray(cov)
        out[f"post/{name}/mean"], out[f"post/{name}/cov"] = np.asarray(d.mean), np.asarray(d.covariance)
    # MarginalModel.step with m = 0, 1, 5, heteroscedastic noise, callable beta (marginal.py:350-368)
    N = 40
    dom = np.sort(rng.random((N, 2)), axis=0)
    out["step/domain"] = dom
    noise = HeteroscedasticNoise([lambda X: 0.05 + 0.1 * X[:, 0], lambda X: 0.2 - 0.1 * X[:, 1]])
    ks = [stationary.Gaussian(lengthscale=0.4), stationary.Matern52(variance=2.0, lengthscale=0.5)]
    distrs = prior_distr(noise, jnp.array(dom), jnp.zeros((0, 2)), jnp.zeros((2, 0)), ks, [ZeroMean()] * 2)
    beta = lambda t: jnp.array([2.0 + 0.1 * t, 3.0])  # noqa: E731
    model = MarginalModel(jr.PRNGKey(1), jnp.array(dom), distrs, beta=beta, noise_rate=noise)
    out["step/noise_at_domain"] = np.asarray(noise.at(jnp.array(dom)))
    for name, oi in (("m1", np.array([7])), ("m5", np.array([3, 30, 12, 12, 25])), ("m0", np.zeros((0,), int)),
                     ("m1b", np.array([39]))):
        y = rng.standard_normal((2, len(oi)))
        model.step(X=model.domain[oi], y=jnp.array(y))
        out[f"step/{name}/obs_idx"], out[f"step/{name}/y"] = oi, y
        for attr in ("means", "variances", "l", "u", "covariances"):
            out[f"step/{name}/{attr}"] = np.asarray(getattr(model, attr))
        out[f"step/{name}/t"] = np.array(model.t)
        out[f"step/{name}/pess"] = np.asarray(model.pessimistic_safe_set)
        out[f"step/{name}/opt"] = np.asarray(model.optimistic_safe_set)
        out[f"step/{name}/pm"] = np.asarray(model.potential_maximizers)
        out[f"step/{name}/pe"] = np.asarray(model.potential_expanders)
        out[f"step/{name}/ucb"] = np.asarray(model.ucb)
        out[f"step/{name}/max_l"] = np.array(float(model.max_l))
        out[f"step/{name}/max_u"] = np.array(float(model.max_u))
    Xs, As = np.array([2, 9, 33]), np.array([1, 9, 20, 38])
    out["step/sqd/X"], out["step/sqd/A"] = Xs, As
    out["step/sqd/value"] = np.asarray(model.sqd_correlation(model.domain[Xs], model.domain[As]))
    # objective_argmax error messages and masking rules (marginal.py:330-348)
    um = UnsafeModel(jr.PRNGKey(2), jnp.array(dom), distrs, beta=jnp.array([2.0, 2.0]), noise_rate=noise)
    msgs = {}
    for name, mdl, F in (
        ("all_neginf", um, np.full(N, -np.inf)),
        ("all_nan", um, np.full(N, np.nan)),
    ):
        try:
            mdl.objective_argmax(jnp.array(F))
            msgs[name] = ""
        except Exception as e:  # noqa: BLE001
            msgs[name] = str(e)
    # a model whose pessimistic safe set is empty: constraint bound far below zero
    d2 = [GaussianDistribution(mean=jnp.array(np.full(N, -5.0)), covariance=jnp.array(K[:N, :N] if K.shape[0] >= N
                                                                                       else np.eye(N)))
          for _ in range(2)]
    sm = MarginalModel(jr.PRNGKey(3), jnp.array(dom), d2, beta=jnp.array([1.0, 1.0]), noise_rate=noise)
    try:
        sm.objective_argmax(jnp.array(rng.random(N)))
        msgs["empty_safe"] = ""
    except Exception as e:  # noqa: BLE001
        msgs["empty_safe"] = str(e)
    for k, v in msgs.items():
        out[f"argmax/msg/{k}"] = np.array(v)
    F = rng.random(N); F[5] = np.nan; F[9] = F.max() if False else 2.0; F[21] = 2.0
    out["argmax/F"] = F
    pt = um.objective_argmax(jnp.array(F))
    out["argmax/choice_idx"] = np.array(int(um.get_indices(pt.reshape(1, -1))[0]))
    out["argmax/tie_set"] = np.flatnonzero(F == np.nanmax(F))
    return out


def next_rows():
    """ContinuousModel.marginalize, LineBO with a fixed sequence of
    directions, the Thompson-sampling masks on preset samples."""
    from lib.algorithms import ts_roi_constructor  # noqa: F401
    from lib.algorithms.line_bo import LineBO, directed_alg_constructor
    from lib.function.unknown import UnknownFunction
    from lib.model.continuous import ContinuousModel
    out = {}
    rng = np.random.default_rng(7)
    # ---- ContinuousModel (lib/model/continuous.py:15-95)
    noise = HomoscedasticNoise(2, jnp.array([0.1, 0.2]))
    ks = [stationary.Gaussian(lengthscale=0.8), stationary.Matern52(variance=1.5, lengthscale=0.6)]
    cm = ContinuousModel(jr.PRNGKey(0), d=2, prior_kernels=ks, beta=jnp.array([2.0, 3.0]), noise_rate=noise)
    Xq = rng.random((15, 2)) * 2 - 1
    m0 = cm.marginalize(jnp.array(Xq))
    out["cont/Xq"] = Xq
    for tag, mm in (("t0", m0),):
        out[f"cont/{tag}/means"], out[f"cont/{tag}/covariances"] = np.asarray(mm.means), np.asarray(mm.covariances)
        out[f"cont/{tag}/l"], out[f"cont/{tag}/u"] = np.asarray(mm.l), np.asarray(mm.u)
    Xa = rng.random((3, 2)) * 2 - 1
    Xb = np.concatenate([rng.random((3, 2)) * 2 - 1, Xa[:1]])  # a repeated location
    ya, yb = rng.standard_normal((2, 3)), rng.standard_normal((2, 4))
    cm.step(jnp.array(Xa), jnp.array(ya))
    cm.step(jnp.array(Xb), jnp.array(yb))
    m7 = cm.marginalize(jnp.array(Xq))
    out["cont/Xa"], out["cont/Xb"], out["cont/ya"], out["cont/yb"] = Xa, Xb, ya, yb
    out["cont/t7/means"], out["cont/t7/covariances"] = np.asarray(m7.means), np.asarray(m7.covariances)
    out["cont/t7/l"], out["cont/t7/u"], out["cont/t7/t"] = np.asarray(m7.l), np.asarray(m7.u), np.array(m7.t)
    out["cont/t7/pess"] = np.asarray(m7.pessimistic_safe_set)
    # ---- LineBO with given directions (lib/algorithms/line_bo.py:31-163)
    dirs = np.array([[1.0, 0.0], [0.3, -0.8], [0.0, 1.0], [-0.6, 0.5], [1.0, 1.0]])

    class FixedLineBO(LineBO):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._it = 0

        def _direction(self, f):
            v = jnp.array(dirs[self._it])
            self._it += 1
            return v

    bounds = np.array([[-1.0, 1.0], [-0.5, 1.5]])
    fun = lambda x: jnp.array([jnp.exp(-jnp.sum((x - jnp.array([0.4, 0.6])) ** 2))])  # noqa: E731
    f = UnknownFunction(q=1, f=fun, use_objective_as_constraint=True)
    noise1 = HomoscedasticNoise(1, jnp.array([0.05]))
    lm = ContinuousModel(jr.PRNGKey(1), d=2, prior_kernels=[stationary.Gaussian(lengthscale=0.5)],
                         beta=jnp.array([2.0]), noise_rate=noise1, use_objective_as_constraint=True)
    x0 = np.array([[0.1, 0.2]])
    lm.step(jnp.array(x0), f.observe(jnp.array(x0)))  # a safe seed (f > 0 there)
    lbo = FixedLineBO(alg=directed_alg_constructor(ITL, pm_roi_constructor), model=lm, bounds=jnp.array(bounds),
                      n=12, eps=0.01)
    out["linebo/dirs"], out["linebo/bounds"], out["linebo/x0"] = dirs, bounds, x0
    for t in range(len(dirs)):
        res = lbo.step(f)
        mm = res["marginal_model"]
        out[f"linebo/{t}/domain"] = np.asarray(mm.domain)
        out[f"linebo/{t}/x"], out[f"linebo/{t}/y"] = np.asarray(res["x"]), np.asarray(res["y"])
        out[f"linebo/{t}/F"], out[f"linebo/{t}/argmax"] = np.asarray(res["F"]), np.asarray(res["argmax"])
        out[f"linebo/{t}/means"], out[f"linebo/{t}/variances"] = np.asarray(mm.means), np.asarray(mm.variances)
        out[f"linebo/{t}/subspace_bounds"] = np.array([float(b) for b in lbo.subspace_bounds()])
    out["linebo/final_X"], out["linebo/final_y"] = np.asarray(lm.X), np.asarray(lm.y)
    out["linebo/encountered_max"] = np.asarray(lbo.encountered_max)
    # ---- TruVar with a fixed ROI (lib/algorithms/tru_var/__init__.py, model.py)
    from lib.algorithms.tru_var import TruVar
    from lib.algorithms.tru_var.model import TruVarModel
    import problems as P
    p3 = P.c3(T=8)
    dom3 = jnp.array(p3["domain"])
    nz = HomoscedasticNoise(1, jnp.array(p3["rho"]))
    d3 = prior_distr(nz, dom3, jnp.array(p3["X0"]), jnp.array(p3["y0"]),
                     [stationary.Matern52(variance=1.0, lengthscale=0.25)], [ZeroMean()])
    tv = TruVarModel(eta=1.2, delta=0.3, r=0.5, key=jr.PRNGKey(4), domain=dom3, distrs=d3, beta=jnp.array([2.0]),
                     noise_rate=nz)
    A3 = np.asarray(p3["roi"])
    alg = TruVar(tv, lambda m: m.domain[A3])
    ftab = TableFunction(dom3, p3["table"], tv.first_constraint)
    Fs, sels, etas = [], [], []
    for t in range(8):
        F = alg.F()
        etas.append(float(tv.eta))
        safe_F = np.asarray(F)
        idx = int(np.flatnonzero(safe_F == np.nanmax(safe_F))[0])
        Fs.append(np.asarray(F)); sels.append(idx)
        X = tv.domain[idx: idx + 1]
        tv.step(X=X, y=ftab.observe(X))
    out["truvar/F"], out["truvar/sel"], out["truvar/eta"] = np.array(Fs), np.array(sels), np.array(etas)
    out["truvar/means"], out["truvar/variances"] = np.asarray(tv.means), np.asarray(tv.variances)
    # ---- Thompson-sampling masks on preset samples (marginal.py:238-312)
    N, q, k = 30, 3, 6
    dom = rng.random((N, 2))
    distrs = [GaussianDistribution(mean=jnp.zeros(N), covariance=jnp.eye(N)) for _ in range(q)]
    tm = MarginalModel(jr.PRNGKey(2), jnp.array(dom), distrs, beta=jnp.array([1.0] * q),
                       noise_rate=HomoscedasticNoise(q, jnp.array([0.1] * q)))
    samples = rng.standard_normal((k, q, N))
    samples[4, 1:] = -1.0  # a sample without any safe point
    out["ts/samples"] = samples

    def preset(which):
        it = iter(range(k))
        tm.sample = lambda k, keys=None: jnp.array(samples[next(it)][:, None, :])  # noqa: E731

    for name, I, J in (("obj0_cons12", jnp.array([0]), jnp.arange(1, 3)),
                       ("obj01_cons2", jnp.array([0, 1]), jnp.array([2])),
                       ("obj0_nocons", jnp.array([0]), None)):
        preset(name)
        out[f"ts/{name}/mask"] = np.asarray(tm.thompson_sampling(k=k, I=I, J=J))
        preset(name)
        out[f"ts/{name}/values"] = np.asarray(tm.thompson_sampling_values(k=k, I=I, J=J))
    return out


def metric_rows():
    """SURVEY 8f-4 remainder: entropy / masked_entropy (marginal.py:113-127, gaussian_distribution.py:141-149),
    information_gain (:151-180), information ratios (lib/metrics/information_ratio.py), per-step
    transductive-learning metrics (lib/metrics/transductive_learning.py)."""
    from lib.metrics.information_ratio import compute_information_ratio, compute_strong_information_ratio
    from lib.metrics.transductive_learning import transductive_learning_metrics_generator
    out = {}
    rng = np.random.default_rng(21)
    N, q = 300, 2
    dom = rng.random((N, 2))
    noise = HomoscedasticNoise(q, jnp.array([0.1, 0.2]))
    X0 = dom[[5, 77, 140]]
    y0 = rng.standard_normal((q, 3))
    distrs = prior_distr(noise, jnp.array(dom), jnp.array(X0), jnp.array(y0),
                         [stationary.Matern52(variance=1.0, lengthscale=0.3), stationary.Gaussian(variance=1.5, lengthscale=0.4)],
                         [ZeroMean()] * q)
    model = MarginalModel(jr.PRNGKey(3), jnp.array(dom), distrs, beta=jnp.array([2.0, 2.0]), noise_rate=noise)
    for t in range(4):
        idx = int(rng.integers(N))
        model.step(X=model.domain[idx: idx + 1], y=jnp.array(rng.standard_normal((q, 1))))
    out["m/domain"], out["m/X0"], out["m/y0"] = dom, X0, y0
    out["m/means"], out["m/covariances"] = np.asarray(model.means), np.asarray(model.covariances)
    mask = np.zeros(N, dtype=bool)
    mask[rng.choice(N, 90, replace=False)] = True
    roi_idx = np.sort(rng.choice(N, 40, replace=False))
    obs_idx = np.sort(rng.choice(N, 70, replace=False))
    out["m/mask"], out["m/roi_idx"], out["m/obs_idx"] = mask, roi_idx, obs_idx
    for jit in (1e-6, 1e-3):
        out[f"m/entropy/{jit}"] = np.asarray(model.entropy(jitter=jit))
        out[f"m/masked_entropy/{jit}"] = np.asarray(model.masked_entropy(jnp.array(mask), jitter=jit))
        out[f"m/distr0_entropy/{jit}"] = np.asarray(model.distr(0).entropy(jitter=jit))
    out["m/total_variance0"] = np.asarray(model.distr(0).total_variance)
    ns = jnp.array(0.05 + 0.2 * rng.random(obs_idx.shape[0]))
    out["m/ig_noise"] = np.asarray(ns)
    out["m/ig/noisy"] = np.asarray(model.distr(0).information_gain(jnp.array(roi_idx), jnp.array(obs_idx), ns))
    out["m/ig/scalar"] = np.asarray(model.distr(1).information_gain(jnp.array(roi_idx), jnp.array(obs_idx), jnp.array(0.1)))
    out["m/ig/overlap"] = np.asarray(model.distr(0).information_gain(jnp.array(roi_idx), jnp.array(roi_idx[:25]), jnp.array(0.3)))
    # (with a sample region smaller than the domain the reference's call fails: it passes the [n] noise
    #  vector next to |B| observation indices, information_ratio.py:51-55; only B = domain is defined)
    out["m/ir/all"] = np.asarray(compute_information_ratio(model, jnp.array(roi_idx), None))
    fm = [0.9, 0.5, float("nan"), 0.7, 0.25]
    out["m/sir/list"] = np.asarray(fm)
    out["m/sir/none"] = np.asarray(compute_strong_information_ratio(fm))
    out["m/sir/prev"] = np.asarray(compute_strong_information_ratio(fm, prev=1.5))
    out["m/sir/prevnan"] = np.asarray(compute_strong_information_ratio(fm, prev=float("nan")))
    mets = transductive_learning_metrics_generator(jnp.array(mask))(model, None, None)
    for k_, v in mets.items():
        out[f"m/tl/{k_}"] = np.asarray(v)
    return out


def run_reference_tests():
    import pytest
    plug = os.path.join("/tmp", "refshim_plugin_dir")
    os.makedirs(plug, exist_ok=True)
    with open(os.path.join(plug, "refshim_plugin.py"), "w") as fh:
        fh.write(f"import sys\nsys.path.insert(0, {HERE!r})\nimport refshim\nrefshim.install()\n"
                 f"sys.path.insert(0, {REF!r})\n")
    sys.path.insert(0, plug)
    tests = [os.path.join(REF, "tests", t) for t in
             ("test_discrete_model.py", "test_continuous_model.py", "test_line_bo.py")]
    rc = pytest.main(["-q", "-p", "refshim_plugin", "-p", "no:cacheprovider", "--rootdir", "/tmp"] + tests)
    assert rc == 0, "the reference's own tests must pass over the stand-in before fixtures are generated"


if __name__ == "__main__":
    stages = sys.argv[1:] or ["tests", "calls", "next", "metrics", "traj"]
    if "tests" in stages:
        run_reference_tests()
    if "calls" in stages:
        calls = single_calls()
        np.savez_compressed(os.path.join(HERE, "reference_calls.npz"), **calls)
        print("wrote", len(calls), "call arrays")
    if "next" in stages:
        nxt = next_rows()
        np.savez_compressed(os.path.join(HERE, "reference_next.npz"), **nxt)
        print("wrote", len(nxt), "next-row arrays")
    if "metrics" in stages:
        met = metric_rows()
        np.savez_compressed(os.path.join(HERE, "reference_metrics.npz"), **met)
        print("wrote", len(met), "metric arrays")
    if "traj" in stages:  # minutes: the stand-in's vmap is a Python loop over all N^2 kernel pairs
        traj = {}
        runs = [(name, None) for name in problems.ALL] + [("c2", "CTL"), ("c4", "MMITL"), ("c1", "VTL"), ("c3", "VTL")]
        for name, rule in runs:
            p = problems.ALL[name]()
            if rule is not None:
                p["T"] = min(p["T"], 8)
            res = run_trajectory(p, rule)
            tag = name if rule is None else f"{name}_{rule}"
            for k, v in res.items():
                traj[f"{tag}/{k}"] = v
            print(tag, "selections", res["sel"].tolist(), "reference choice", res["ref_choice"].tolist(),
                  "ties", res["n_ties"].tolist(), flush=True)
        np.savez_compressed(os.path.join(HERE, "reference_trajectories.npz"), **traj)
        print("wrote", len(traj), "trajectory arrays")
