"""Committed fixtures (tests/golden/oracle_trajectories.npz): the oracle must
still reproduce them (CPU), and the CUDA path must match them without the
oracle in the loop (GPU, teacher-forced on the fixtures' selections)."""
import os

import numpy as np
import pytest

import problems

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_trajectories.npz"))
TIE_TOL = 1e-9


@pytest.mark.parametrize("name", sorted(problems.ALL))
def test_oracle_reproduces_fixtures(name):
    p = problems.ALL[name]()
    model, alg, f = problems.build_oracle(p)
    for t in range(p["T"]):
        F = alg.masked_F()
        fin = np.isfinite(F)
        np.testing.assert_allclose(F[fin], FIX[f"{name}/F"][t][fin], rtol=1e-7, atol=1e-10)
        idx = int(FIX[f"{name}/sel"][t])
        assert F[idx] >= np.nanmax(np.where(model.operating_safe_set, F, -np.inf)) - TIE_TOL
        X = model.domain[idx: idx + 1]
        model.step(X, f.observe(X))
    np.testing.assert_allclose(model.means, FIX[f"{name}/means"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(model.variances, FIX[f"{name}/variances"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(problems.ALL))
def test_cuda_path_matches_fixtures(cuda, name):
    import torch
    p = problems.ALL[name]()
    model, alg, f = problems.build_cuda(p)
    agree = 0
    for t in range(p["T"]):
        F = alg.F()
        idx_c = model._argmax_index(F, alg.sample_mask())
        Fg = FIX[f"{name}/F"][t]
        Fc = F.cpu().numpy()
        fin = np.isfinite(Fg)
        np.testing.assert_allclose(Fc[fin], Fg[fin], rtol=1e-6, atol=1e-9)
        assert np.array_equal(np.isneginf(Fc), np.isneginf(Fg))
        safe = model.operating_safe_set.cpu().numpy()
        mx = np.nanmax(np.where(safe, Fg, -np.inf))
        assert Fg[idx_c] >= mx - TIE_TOL * max(1.0, abs(mx))  # the selection is a fixture maximiser
        idx = int(FIX[f"{name}/sel"][t])
        agree += int(idx == idx_c)
        X = model.domain[idx: idx + 1]
        model.step(X, f.observe(X))
    assert agree >= p["T"] - 1
    tol = dict(rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(model.means.cpu().numpy(), FIX[f"{name}/means"], **tol)
    np.testing.assert_allclose(model.variances.cpu().numpy(), FIX[f"{name}/variances"], **tol)
    np.testing.assert_allclose(model.l.cpu().numpy(), FIX[f"{name}/l"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(model.u.cpu().numpy(), FIX[f"{name}/u"], rtol=1e-9, atol=1e-10)
    assert torch.equal(model.variances, torch.diagonal(model.covariances, dim1=1, dim2=2))
