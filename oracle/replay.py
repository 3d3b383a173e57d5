"""Low-rank replay oracle: the reference's posterior at FULL domain size without an N x N matrix.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference's `MarginalModel.step` (lib/model/marginal.py:350-368 ->
lib/gp/gaussian_distribution.py:16-37 with m = 1) rewrites the whole N x N covariance after every
observation:  Sigma' = Sigma - k^T w,  k = Sigma[x, :],  w = (k / L) / L,  L = sqrt(k[x] + rho^2).
Entry (r, c) after t observations is therefore

    Sigma_t[r, c] = ((K[r, c] - k_0[r] w_0[c]) - k_1[r] w_1[c]) - ... - k_{t-1}[r] w_{t-1}[c]

with the SAME sequence of rounded operations as the reference (NumPy does not fuse the multiply into
the subtraction).  Only the observed rows are ever needed to continue:

    k_i = K[x_i, :] - sum_{j<i} k_j[x_i] w_j[:]          O(i N) per observation

so means, variances, confidence bounds over all N candidates cost O(t^2 N) on the host and any block
Sigma_t[R, C] costs O(t |R| |C|).  `itl_directed` evaluates the reference's directed-ITL objective
(lib/algorithms/itl.py:33-60) on a subset C of candidates from Sigma_t[A, A] and Sigma_t[A, C]: the
same formula (Cholesky of Sigma_AA + jitter^2 I, (L L^T)^-1 Sigma_AC, diag of the Schur complement),
restricted to the columns asked for.  This is what the full-size parity tests check the CUDA path
against (tests/test_gpu_replay.py, tools/mp_parity.py); at small N it is itself checked against the
dense restatement in ref_numpy.py (tests/test_oracle.py)."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import scipy.linalg as sla

from .ref_numpy import DEFAULT_JITTER, Kernel


class LowRankReplay:
    def __init__(self, kernel: Kernel, domain: np.ndarray, rho: float, beta: float, mean0: float = 0.0):
        self.kernel = kernel
        self.domain = np.asarray(domain, dtype=np.float64)
        self.N = self.domain.shape[0]
        self.rho = np.full(self.N, float(rho))
        self.beta = float(beta)
        self.k: list = []   # observed rows k_i [N]
        self.w: list = []   # weights w_i [N]
        self.x: list = []
        self.mean = np.full(self.N, float(mean0))
        diff0 = np.zeros((1, self.domain.shape[1]))
        self.var = np.full(self.N, float(kernel.pair(diff0)[0]))  # stationary kernel: K[c, c] = k(0)
        sd = np.sqrt(self.var)
        self.l = self.mean - self.beta * sd   # marginal.py:160-164
        self.u = self.mean + self.beta * sd

    # Gram entries (lib/gp/kernels/base.py:17-25): K[r, c] = k(domain[r], domain[c])
    def gram(self, R: Sequence[int], C: Optional[Sequence[int]] = None) -> np.ndarray:
        Y = self.domain if C is None else self.domain[np.asarray(C)]
        return self.kernel.cross_covariance(Y, self.domain[np.asarray(R)])  # [len(R), len(Y)] (base.py:20-22)

    def row(self, x: int) -> np.ndarray:
        """Sigma_t[x, :] (all N columns)."""
        r = self.gram([x])[0]
        for kj, wj in zip(self.k, self.w):
            r = r - kj[x] * wj
        return r

    def block(self, R: Sequence[int], C: Sequence[int]) -> np.ndarray:
        """Sigma_t[R, C]."""
        R, C = np.asarray(R), np.asarray(C)
        B = self.gram(R, C)
        for kj, wj in zip(self.k, self.w):
            B = B - np.outer(kj[R], wj[C])
        return B

    def observe(self, x: int, y: float) -> None:
        """One `MarginalModel.step` with m = 1 at domain index x."""
        k = self.row(x)
        s = k[x] + self.rho[x] ** 2          # gaussian_distribution.py:29-32
        L = np.sqrt(s)
        w = (k / L) / L                       # utils.py:22-23 for a 1 x 1 system
        self.mean = self.mean + w * (y - self.mean[x])   # :120
        self.var = self.var - k * w                        # diag of :36
        sd = np.sqrt(self.var)
        self.l = np.maximum(self.l, self.mean - self.beta * sd)   # marginal.py:366-367
        self.u = np.minimum(self.u, self.mean + self.beta * sd)
        self.k.append(k); self.w.append(w); self.x.append(int(x))

    def itl_directed(self, A: Sequence[int], C: Sequence[int], jitter: float = DEFAULT_JITTER) -> np.ndarray:
        """F[c] for c in C (itl.py:41-58): 0.5 max(log((var + rho^2) / (var_{c|A} + rho^2)), 0) with
        var_{c|A} = var[c] - Sigma[A, c]^T (Sigma[A, A] + jitter^2 I)^-1 Sigma[A, c]."""
        A, C = np.asarray(A), np.asarray(C)
        S = self.block(A, A) + np.eye(len(A)) * jitter ** 2      # utils.py:44-50: the default is squared too
        B = self.block(A, C)
        L = np.linalg.cholesky(S)
        Z = sla.solve_triangular(L, B, lower=True)               # (the reference: LU solves on L, L^T; utils.py:23)
        W = sla.solve_triangular(L.T, Z, lower=False)
        post = self.var[C] - np.einsum("ac,ac->c", B, W)         # diag(Sigma - Kxt^T W) (gaussian_distribution.py:36)
        nz = self.rho[C] ** 2
        return 0.5 * np.maximum(np.log((self.var[C] + nz) / (post + nz)), 0.0)
