"""Debug: c1 (1-D safe BO, ITL-PM, RBF) trajectory; per step the conditioning kernels by hand, NaN checks."""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transductive-active-learning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

def child(tag):
    import torch
    import oracle as O
    from test_gpu_e2e import _build_pair, _bumps
    from lib.algorithms import pm_roi_constructor
    from lib.algorithms.itl import ITL
    from lib.gp.kernels import stationary
    from talgp import _abi
    N = 400
    domain = np.linspace(-2, 8, N)[:, None]
    rng = np.random.default_rng(8)
    table = (_bumps(domain[:, 0]) + 0.1 * rng.standard_normal(N))[None, :]
    seed_idx = np.sort(rng.choice(np.where((domain[:, 0] > 4.5) & (domain[:, 0] < 6))[0], 8))
    X0, y0 = domain[seed_idx], table[:, seed_idx]
    mo, mc = _build_pair(domain, [O.Gaussian(lengthscale=1.0)], [stationary.Gaussian(lengthscale=1.0)],
                         beta=[3.0], rho=[0.1], X0=X0, y0=y0, fc_obj=True)
    alg_o = O.ITL(mo, O.pm_roi_constructor)
    alg_c = ITL(mc, pm_roi_constructor)
    gp = mc._gp
    lib = gp.lib
    p = lambda t: C.c_void_p(t.data_ptr())
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for t in range(25):
        roi = alg_c.roi
        m = roi.m
        m_pad, ncols = gp._cond_dims(m)
        S = torch.empty((m_pad, m_pad), dtype=torch.float64, device="cuda")
        B = torch.empty((m_pad, ncols), dtype=torch.float64, device="cuda")
        work = torch.zeros((int(lib.tal_potrf_work_bytes(m_pad)) // 8,), dtype=torch.float64, device="cuda")
        gp.flush()
        _abi.check(lib.tal_gather_roi(p(gp._cov[0]), gp.ld, gp.N, p(roi.idx), m, m_pad, 1e-5, p(S), m_pad, p(B), ncols, gp.dt, st()))
        S_in = S.clone()
        _abi.check(lib.tal_potrf_lower(p(S), m_pad, m_pad, p(work), st()))
        torch.cuda.synchronize()
        L = torch.tril(S)
        nanL = int(torch.isnan(L).sum())
        Fo = alg_o.masked_F()
        subset = mc.safe_set_is_subset_of(A=roi)
        Fc = alg_c.F()
        nanF = int(torch.isnan(Fc).sum())
        w = np.linalg.eigvalsh(S_in.cpu().numpy()[:m, :m])
        print(tag, "step", t, "m", m, "m_pad", m_pad, "undirected" if subset else "directed", "nan(L)", nanL, "nan(F)", nanF,
              "eig min %.3e" % w[0], "max|Fc-Fo| %.2e" % np.nanmax(np.abs(np.where(np.isfinite(Fo), Fc.cpu().numpy() - Fo, 0))), flush=True)
        if nanL or nanF:
            np.save(os.path.join(ROOT, "gpurun_out", f"dbg_c1_Sin_{tag}.npy"), S_in.cpu().numpy())
            np.save(os.path.join(ROOT, "gpurun_out", f"dbg_c1_L_{tag}.npy"), S.cpu().numpy())
            np.save(os.path.join(ROOT, "gpurun_out", f"dbg_c1_F_{tag}.npy"), Fc.cpu().numpy())
            r = torch.nonzero(torch.isnan(L))
            if r.numel():
                print("   first NaN in L at", r[0].tolist())
            break
        idx = int(mo.objective_argmax_index(Fo))
        X = domain[idx: idx + 1]
        y = table[:, [idx]]
        mo.step(X, y); mc.step(X, y)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1]); sys.exit(0)
    for tag, env in (("new", "1"), ("old", "0")):
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=dict(os.environ, TAL_POTRF_DIAG=env))
