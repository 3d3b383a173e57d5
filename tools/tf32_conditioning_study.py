"""Would a TF32 (tcgen05 kind::tf32) triangular solve do for the target-set conditioning of directed ITL
(lib/algorithms/itl.py:41-48, lib/utils.py:22-23) in the fp32-storage mode?  Numerical study on the CPU at
the C3 geometry (16^4 jittered grid, Matern-5/2 l = 0.25, m = 4,096 targets, jitter^2 = 1e-10):

    V = L^-1 Sigma_A. ,  cs[x] = ||V[:, x]||^2 ,  F[x] = 0.5 log((var + rho^2) / (var - cs + rho^2))

with the m x m Cholesky in fp64 (as SURVEY.md 7 proposes) and the m x N solve emulated in
  fp64                      the shipped DMMA path
  fp32                      what a 3xTF32 split (error-compensated tensor-core product) would give at best
  tf32                      operands rounded to 10 mantissa bits, fp32 accumulation: tcgen05.mma kind::tf32
The solve is the blocked one of csrc/condvar.cu (128-row blocks, explicit inverses of the diagonal blocks).
Prints the condition number of Sigma_AA + jitter^2 I, the error of cs and F over a sample of candidates,
and how many candidates could be mis-ranked against the best one.  Output committed as
profiles/r02_tf32_conditioning_study.txt.
"""
import sys
import numpy as np
import scipy.linalg as sl

NB = 128


def tf32(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = (u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def matern52(X, Y, l):
    d = np.sqrt(np.maximum(((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1), 0.0)) / l
    s5 = np.sqrt(5.0) * d
    return (1.0 + s5 + 5.0 / 3.0 * d * d) * np.exp(-s5)


def blocked_trsm(L, B, mode):
    """X = L^-1 B by block rows; products in `mode` arithmetic, storage of X in the mode's type."""
    m = L.shape[0]
    nblk = m // NB
    Linv = [np.linalg.inv(L[i * NB:(i + 1) * NB, i * NB:(i + 1) * NB]) for i in range(nblk)]
    if mode == "fp64":
        X = B.astype(np.float64).copy()
        mm = lambda a, b: a @ b
    else:
        X = B.astype(np.float32).copy()
        if mode == "fp32":
            mm = lambda a, b: a.astype(np.float32) @ b.astype(np.float32)
        else:
            mm = lambda a, b: tf32(a) @ tf32(b)
    for i in range(nblk):
        r = slice(i * NB, (i + 1) * NB)
        if i > 0:
            X[r] = X[r] - mm(L[r, : i * NB], X[: i * NB])
        X[r] = mm(Linv[i], X[r])
    return X.astype(np.float64)


def main():
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n_s = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    g, d, N, l, rho = 16, 4, 65536, 0.25, 0.1
    rng = np.random.default_rng(0)
    axes = np.meshgrid(*([np.arange(g) / g] * d), indexing="ij")
    domain = np.stack(axes, -1).reshape(-1, d) + rng.random((N, d)) / g
    A = np.sort(rng.choice(N, m, replace=False))
    cand = np.sort(np.random.default_rng(5).choice(N, n_s, replace=False))
    XA = domain[A]
    S = np.concatenate([matern52(XA[i:i + 512], XA, l) for i in range(0, m, 512)])
    B = np.concatenate([matern52(XA[i:i + 512], domain[cand], l) for i in range(0, m, 512)])
    w = np.linalg.eigvalsh(S)
    print(f"Sigma_AA: m = {m}, eigenvalues in [{w[0]:.3e}, {w[-1]:.3e}], cond(Sigma_AA + 1e-10 I) = {(w[-1] + 1e-10) / (w[0] + 1e-10):.3e}")
    print(f"fp32 storage of Sigma_AA alone perturbs it by ~ eps32 * lambda_max = {6e-8 * w[-1]:.2e} "
          f"({'ABOVE' if 6e-8 * w[-1] > w[0] + 1e-10 else 'below'} lambda_min + jitter^2 = {w[0] + 1e-10:.2e})")
    L = np.linalg.cholesky(S + 1e-10 * np.eye(m))
    var = np.ones(n_s)  # prior: k(x, x) = 1
    ref = None
    for mode in ("fp64", "fp32", "tf32"):
        V = blocked_trsm(L, B, mode)
        cs = (V * V).sum(0)
        F = 0.5 * np.maximum(np.log((var + rho ** 2) / (var - cs + rho ** 2)), 0.0)
        if ref is None:
            ref = (cs, F)
            exact = sl.solve_triangular(L, B, lower=True)
            cse = (exact * exact).sum(0)
            print(f"fp64 blocked solve vs LAPACK trsm: max |cs - cs_lapack| / cs = {np.max(np.abs(cs - cse) / cse):.2e}")
            order = np.argsort(-F)
            print(f"F over the sample: max {F[order[0]]:.6f}, runner-up {F[order[1]]:.6f} (gap {F[order[0]] - F[order[1]]:.2e}), "
                  f"median {np.median(F):.4f}; conditional variance var - cs in [{(var - cs).min():.3e}, {(var - cs).max():.3e}]")
            continue
        cs0, F0 = ref
        err_cs = np.abs(cs - cs0)
        err_F = np.abs(F - F0)
        neg = int(((var - cs) < 0).sum())
        best = np.argmax(F0)
        misrank = int((F > F[best]).sum())
        print(f"{mode}: max |cs err| = {err_cs.max():.3e} (relative to the conditional variance: {np.max(err_cs / np.maximum(var - cs0, 1e-300)):.3e}), "
              f"max |F err| = {err_F.max():.3e}, median |F err| = {np.median(err_F):.3e}, negative conditional variances: {neg}, "
              f"candidates ranked above the true maximiser: {misrank}, arg-max unchanged: {bool(np.argmax(F) == best)}")


if __name__ == "__main__":
    main()
