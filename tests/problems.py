"""Small seeded problems shared by the fixture generator and the tests: scaled-
down versions of BASELINE.json's configs, as plain NumPy inputs."""
import numpy as np


def _bumps(x):
    return (np.exp(-(x - 1) ** 2) + 1.5 * np.exp(-(x - 5) ** 2 / 0.5) + 0.8 * np.exp(-(x - 6.5) ** 2 / 0.2)
            - 0.3 - 0.05 * (x - 3) ** 2 * (x < 0))


def c1(N=300, T=20):
    """1-D safe BO, ITL with the potential-maximizer ROI, RBF, objective as constraint."""
    domain = np.linspace(-2, 8, N)[:, None]
    rng = np.random.default_rng(8)
    table = (_bumps(domain[:, 0]) + 0.1 * rng.standard_normal(N))[None, :]
    seed_idx = np.sort(rng.choice(np.where((domain[:, 0] > 4.5) & (domain[:, 0] < 6))[0], 8))
    return dict(name="c1", domain=domain, table=table, X0=domain[seed_idx], y0=table[:, seed_idx],
                kernels=[("Gaussian", 1.0, 1.0)], beta=[3.0], rho=[0.1], fc_obj=True, rule="ITL", roi="pm", T=T)


def c2(T=12, step=0.3):
    """2-D safe BO, objective + 1 constraint GP, VTL with the potential-maximizer ROI."""
    g = np.arange(-3, 3, step)
    domain = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    N = len(domain)
    rng = np.random.default_rng(3)
    f = 2.0 * np.exp(-0.5 * np.sum((domain - 0.5) ** 2, axis=1))
    gfun = 1.5 * np.exp(-0.25 * np.sum(domain ** 2, axis=1)) - 0.4
    table = np.stack([f, gfun]) + 0.1 * rng.standard_normal((2, N))
    seed = np.array([int(np.argmin(np.sum(domain ** 2, axis=1)))])
    return dict(name="c2", domain=domain, table=table, X0=domain[seed], y0=table[:, seed],
                kernels=[("Gaussian", 1.0, 1.0)] * 2, beta=[3.0, 3.0], rho=[0.1, 0.1], fc_obj=False,
                rule="VTL", roi="pm", T=T)


def c3(T=12):
    """Transductive ITL with a fixed target set, d = 4 Matérn-5/2, unconstrained."""
    g = np.arange(5) / 5.0
    rng = np.random.default_rng(0)
    domain = np.stack(np.meshgrid(g, g, g, g, indexing="ij"), -1).reshape(-1, 4)
    domain = domain + rng.random(domain.shape) / 5
    N = len(domain)
    A = np.sort(rng.choice(N, 80, replace=False))
    table = (np.sin(3 * domain.sum(axis=1)) + 0.1 * np.random.default_rng(1).standard_normal(N))[None, :]
    return dict(name="c3", domain=domain, table=table, X0=np.zeros((0, 4)), y0=np.zeros((1, 0)),
                kernels=[("Matern52", 1.0, 0.25)], beta=[1.0], rho=[0.1], fc_obj=False, rule="ITL", roi=A, T=T)


def c4(T=10):
    """Batched safe BO: objective + 3 constraint GPs, VTL-PM (a scaled-down config 4)."""
    rng = np.random.default_rng(11)
    domain = rng.random((400, 3))
    N = len(domain)
    z = rng.random((3, 3))
    f = np.sin(4 * domain.sum(axis=1))
    cons = np.stack([0.9 - np.sum((domain - z[i]) ** 2, axis=1) for i in range(3)])
    table = np.concatenate([f[None], cons]) + 0.05 * rng.standard_normal((4, N))
    seed = np.argsort(-cons.min(axis=0))[:5]
    seed = np.sort(seed)
    return dict(name="c4", domain=domain, table=table, X0=domain[seed], y0=table[:, seed],
                kernels=[("Matern52", 1.0, 0.3), ("Matern52", 1.0, 0.4), ("Gaussian", 1.0, 0.4), ("Matern32", 1.0, 0.5)],
                beta=[2.0] * 4, rho=[0.05] * 4, fc_obj=False, rule="VTL", roi="pm", T=T)


ALL = {"c1": c1, "c2": c2, "c3": c3, "c4": c4}


def build_oracle(p, rule=None):
    import oracle as O
    q = len(p["kernels"])
    ks = [getattr(O, k)(variance=v, lengthscale=l) for k, v, l in p["kernels"]]
    noise = O.HomoscedasticNoise(q, np.asarray(p["rho"]))
    distrs = O.prior_distr(noise, p["domain"], p["X0"], p["y0"], ks, [O.ZeroMean()] * q)
    model = O.MarginalModel(0, p["domain"], distrs, beta=np.asarray(p["beta"], float), noise_rate=noise,
                            use_objective_as_constraint=p["fc_obj"])
    if isinstance(p["roi"], str):
        ctor = {"pm": O.pm_roi_constructor, "pe": O.pe_roi_constructor}[p["roi"]]
    else:
        A = p["roi"]
        ctor = lambda m: m.domain[A]
    alg = {"ITL": O.ITL, "VTL": O.VTL, "CTL": O.CTL, "MMITL": O.MMITL}[rule or p["rule"]](model, ctor)
    return model, alg, O.TableFunction(p["domain"], p["table"])


def build_cuda(p, comm=None, rule=None, storage="full", **alg_kw):
    from lib.algorithms import index_roi_constructor, pe_roi_constructor, pm_roi_constructor
    from lib.algorithms.ctl import CTL
    from lib.algorithms.itl import ITL
    from lib.algorithms.mm_itl import MMITL
    from lib.algorithms.vtl import VTL
    from lib.function import TableFunction
    from lib.gp.kernels import stationary
    from lib.gp.means import ZeroMean
    from lib.gp.prior import prior_distr
    from lib.model.marginal import MarginalModel
    from lib.noise import HomoscedasticNoise
    q = len(p["kernels"])
    ks = [getattr(stationary, k)(variance=v, lengthscale=l) for k, v, l in p["kernels"]]
    noise = HomoscedasticNoise(q, np.asarray(p["rho"]))
    distrs = prior_distr(noise, p["domain"], p["X0"], p["y0"], ks, [ZeroMean()] * q, comm=comm, storage=storage)
    model = MarginalModel(0, p["domain"], distrs, beta=np.asarray(p["beta"], float), noise_rate=noise,
                          use_objective_as_constraint=p["fc_obj"])
    if isinstance(p["roi"], str):
        ctor = {"pm": pm_roi_constructor, "pe": pe_roi_constructor}[p["roi"]]
    else:
        ctor = index_roi_constructor(p["roi"])
    alg = {"ITL": ITL, "VTL": VTL, "CTL": CTL, "MMITL": MMITL}[rule or p["rule"]](model, ctor, **alg_kw)
    return model, alg, TableFunction(model.domain, p["table"], first_constraint=model.first_constraint)
