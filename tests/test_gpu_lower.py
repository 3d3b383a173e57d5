"""Lower-block ("symmetric") storage of the posterior covariance (include/tal_b200.h):
only the blocks on or below the diagonal are maintained, every reader of a full row
goes through the mirrored entries.  It must reproduce the full-storage engine — the
same selections, means / variances / bounds to 1e-12 (the two differ only by the
rounding of k_a w_b versus k_b w_a in the mirrored entries) — and the reference
fixtures at the usual tolerances."""
import os

import numpy as np
import pytest
import torch

import problems

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _pair(cuda, N, q, dtype, seed=0):
    from talgp import _abi
    from talgp.engine import DeviceGP
    rng = np.random.default_rng(seed)
    dom = torch.as_tensor(rng.random((N, 3)), device=cuda)
    gps = []
    for storage in ("full", "lower"):
        gp = DeviceGP(N, q, dtype=dtype, device=cuda, storage=storage)
        for i in range(q):
            gp.build_gram(i, _abi.KERNEL_MATERN52 if i % 2 == 0 else _abi.KERNEL_GAUSSIAN, 1.0 + 0.5 * i,
                          0.3 + 0.1 * i, dom)
        gp.rho.fill_(0.1)
        gp.init_bounds([2.0] * q)
        gps.append(gp)
    return gps, rng


@pytest.mark.parametrize("dtype,N", [(torch.float64, 2600), (torch.float64, 1024), (torch.float64, 333),
                                     (torch.float32, 4500)])
def test_rank1_lower_matches_full(cuda, dtype, N):
    q = 2
    (full, low), rng = _pair(cuda, N, q, dtype)
    blk = int(low.lib.tal_sym_block(low.dt))
    for t in range(12):
        idx = int(rng.integers(N))
        y = rng.standard_normal(q)
        full.observe(idx, y, [2.0] * q)
        low.observe(idx, y, [2.0] * q)
        if t == 5:  # a reader of full rows in the middle: mirror, then keep going
            low.ensure_full()
    tol = dict(rtol=1e-12, atol=1e-14) if dtype == torch.float64 else dict(rtol=1e-3, atol=5e-5)  # two fp32 roundings of a cancelling update
    for name in ("mean", "var", "l", "u"):
        torch.testing.assert_close(getattr(low, name), getattr(full, name), **tol)
    # the kept blocks carry the diagonal: var == diag(cov) bit for bit, as in full storage
    low.flush()  # (updates may still be pending: they are applied in blocks)
    assert low._upper_stale
    for i in range(q):
        assert torch.equal(torch.diagonal(low.cov[i, :, :N]).to(torch.float64), low.var[i])
    # blocks above the diagonal were not touched since the mirror at t == 5 ...
    r = torch.arange(N, device=cuda)
    cend = torch.clamp((r // blk + 1) * blk, max=low.ld)
    kept = torch.arange(low.ld, device=cuda)[None, :] < cend[:, None]
    # ... and after a mirror the whole matrix agrees with full storage
    cov = low.covariances_view()
    assert not low._upper_stale
    torch.testing.assert_close(cov, full.covariances_view(), **tol)
    up = ~kept[:, :N]  # the mirrored entries are exact copies of their images
    assert torch.equal(cov[:, up], cov.transpose(1, 2)[:, up])


def test_extract_row_lower_is_the_mirrored_row(cuda):
    (full, low), rng = _pair(cuda, 2100, 1, torch.float64, seed=3)
    for idx in (0, 5, 1023, 1024, 1500, 2099):
        low.observe(idx, [0.3], [2.0])
        full.observe(idx, [0.3], [2.0])
    for idx in (7, 1023, 1024, 2050):
        low.extract_row(None, idx)
        full.extract_row(None, idx)
        torch.testing.assert_close(low.kbuf, full.kbuf, rtol=1e-12, atol=1e-14)
        assert float(low.kbuf[0, low.N:].abs().sum()) == 0.0  # pad columns stay zero


def test_general_m_conditioning_lower(cuda):
    (full, low), rng = _pair(cuda, 2300, 1, torch.float64, seed=5)
    oi = torch.as_tensor(rng.choice(2300, 70, replace=False), device=cuda)
    y = torch.as_tensor(rng.standard_normal(70), device=cuda)
    ns = torch.full((70,), 0.1, device=cuda, dtype=torch.float64)
    for gp in (full, low):
        gp.condition(0, oi, y, ns)
    torch.testing.assert_close(low.mean, full.mean, rtol=1e-11, atol=1e-13)
    torch.testing.assert_close(low.var, full.var, rtol=1e-11, atol=1e-13)
    torch.testing.assert_close(low.covariances_view(), full.covariances_view(), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("name,rule", [("c1", None), ("c2", None), ("c3", None), ("c4", None), ("c2", "CTL"),
                                       ("c4", "MMITL"), ("c3", "VTL")])
def test_lower_storage_matches_reference_trajectories(cuda, name, rule):
    """The drop-in API over a lower-block model against the golden vectors made by
    the reference's own source (tests/golden/reference_trajectories.npz)."""
    TRAJ = np.load(os.path.join(HERE, "golden", "reference_trajectories.npz"))
    tag = name if rule is None else f"{name}_{rule}"
    p = problems.ALL[name]()
    T = int(TRAJ[f"{tag}/sel"].shape[0])
    model, alg, f = problems.build_cuda(p, rule=rule, storage="lower")
    assert model._gp.lower
    ftol = dict(rtol=1e-6, atol=1e-9) if (rule or p["rule"]) == "ITL" else dict(rtol=1e-9, atol=1e-11)
    for t in range(T):
        F = alg.F()
        idx_sel = model._argmax_index(F, alg.sample_mask())
        Fn, Fg = F.cpu().numpy(), TRAJ[f"{tag}/F"][t]
        fin = np.isfinite(Fg)
        assert np.array_equal(np.isneginf(Fn), np.isneginf(Fg))
        np.testing.assert_allclose(Fn[fin], Fg[fin], err_msg=f"step {t}", **ftol)
        safe = model.operating_safe_set.cpu().numpy()
        mx = np.nanmax(np.where(safe, Fg, -np.inf))
        assert Fg[idx_sel] >= mx - 1e-9 * max(1.0, abs(mx))
        idx = int(TRAJ[f"{tag}/sel"][t])
        X = model.domain[idx: idx + 1]
        model.step(X, f.observe(X))
    for attr in ("means", "variances", "l", "u"):
        np.testing.assert_allclose(getattr(model, attr).cpu().numpy(), TRAJ[f"{tag}/{attr}"], err_msg=attr,
                                   rtol=1e-10, atol=1e-11)
    model._gp.flush()
    assert model._gp._upper_stale  # no scoring rule needed the mirrored blocks
    cov = model.covariances
    torch.testing.assert_close(torch.diagonal(cov, dim1=1, dim2=2), model.variances, rtol=0, atol=0)


def test_lower_fullsize_properties_n65536(cuda):
    """BASELINE configs[2] in lower-block storage at full size: the properties that do not
    need an oracle — var == diag(cov) bit for bit after the updates, the observed points'
    variance collapses towards rho^2-level, selections equal the full-storage run."""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from bench import WORKLOADS, make_problem
    from talgp import _abi
    from talgp.engine import DeviceGP
    w = WORKLOADS["c3"]
    domain, A, table = make_problem(w)
    N = w["N"]
    dom = torch.as_tensor(domain, device=cuda)
    roi = torch.as_tensor(A, device=cuda)
    tab = torch.as_tensor(table, device=cuda)
    sels = {}
    for storage in ("lower", "full"):
        gp = DeviceGP(N, 1, device=cuda, storage=storage)
        gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, w["lengthscale"], dom)
        gp.rho.fill_(w["rho"]); gp.init_bounds([1.0])
        sel = []
        for t in range(6):
            F = gp.score_itl_directed(roi, len(A), incremental=True, roi_key=roi)
            rec = gp.masked_argmax(F, None, 1, None)
            sel.append(int(rec[0].item()))
            gp.observe(rec[0:1].clone(), tab, [1.0], y_gather=True, y_stride=N)
        sels[storage] = sel
        assert torch.equal(torch.diagonal(gp.cov[0, :, :N]), gp.var[0])
        assert float(gp.var[0, sel].max()) < 0.02 and float(gp.var.min()) > 0
        if storage == "lower":
            v_low = gp.var.clone()
        else:
            torch.testing.assert_close(v_low, gp.var, rtol=1e-11, atol=1e-14)
        del gp
        torch.cuda.empty_cache()
    assert sels["lower"] == sels["full"]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("frac", [0.3, 1.0])
def test_dense_roi_row_reduce_lower_matches_full(cuda, dtype, frac):
    """ROIs that cover a large part of the domain (potential maximisers: m ~ N) take the streamed + masked
    form of the mirrored column sums (tal_colsum_lower_dense); same scores as full storage for all four rules,
    with the mask supplied (ROI.from_mask) or rebuilt from the index list."""
    from talgp import _abi
    N = 2100
    (full, low), rng = _pair(cuda, N, 2, dtype, seed=5)
    ys = rng.standard_normal((3, 2))
    for g in (full, low):  # a few observations so that the matrices are not the bare Gram matrices
        for t, idx in enumerate((7, 900, 2050)):
            g.observe(idx, ys[t], [2.0, 2.0])
    m = int(frac * N)
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
    mask = torch.zeros((N,), dtype=torch.uint8, device=cuda)
    mask[roi] = 1
    tol = dict(rtol=1e-11, atol=1e-12) if dtype == torch.float64 else dict(rtol=2e-5, atol=1e-6)
    for rule, kw in ((_abi.RULE_VTL, {}), (_abi.RULE_CTL, {}), (_abi.RULE_MMITL, {}),
                     (_abi.RULE_TRUVAR, dict(p0=[1.5, 1.5], p1=0.01))):
        Ff = full.score_rows(rule, roi, m, **kw)
        Fl = low.score_rows(rule, roi, m, roi_mask=mask, **kw)
        Fl2 = low.score_rows(rule, roi, m, **kw)
        torch.testing.assert_close(Fl, Ff, **tol)
        assert torch.equal(Fl, Fl2)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("arith", ["fma", "two_rounding"])
def test_fused_update_and_dense_row_reduce(cuda, dtype, arith, monkeypatch):
    """A VTL-type step in lower storage with a large ROI: the ONE pending rank-1 update and both halves of the next
    score in one pass over the matrix (tal_update_colsum_lower_dense).  Same scores as update-then-score (the sums
    run over the same updated values; only the summation order of the partials differs), and the matrix is
    bit-identical to the one the separate update pass produces."""
    from talgp import _abi
    from talgp.engine import DeviceGP
    N, q = 2100, 2
    rng = np.random.default_rng(9)
    dom = torch.as_tensor(rng.random((N, 3)), device=cuda)

    def make():
        gp = DeviceGP(N, q, dtype=dtype, device=cuda, storage="lower", arith=arith)
        for i in range(q):
            gp.build_gram(i, _abi.KERNEL_MATERN52 if i == 0 else _abi.KERNEL_GAUSSIAN, 1.0 + 0.5 * i, 0.3 + 0.1 * i, dom)
        gp.rho.fill_(0.1)
        gp.init_bounds([2.0] * q)
        return gp

    fused, plain = make(), make()
    m = 1700
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
    tol = dict(rtol=1e-11, atol=1e-12) if dtype == torch.float64 else dict(rtol=2e-5, atol=1e-6)
    rules = ((_abi.RULE_VTL, {}), (_abi.RULE_CTL, {}), (_abi.RULE_MMITL, {}), (_abi.RULE_TRUVAR, dict(p0=[1.5, 1.5], p1=0.01)))
    for step, (rule, kw) in enumerate(rules * 2):
        idx = int(rng.integers(N))
        y = rng.standard_normal(q)
        for g in (fused, plain):
            g.observe(idx, y, [2.0] * q)
        assert fused._pending == 1 and plain._pending == 1
        monkeypatch.setenv("TAL_FUSED_VTL", "1")
        Ff = fused.score_rows(rule, roi, m, **kw)
        assert fused._pending == 0
        monkeypatch.setenv("TAL_FUSED_VTL", "0")
        Fp = plain.score_rows(rule, roi, m, **kw)
        torch.testing.assert_close(Ff, Fp, **tol)
        blk = int(fused.lib.tal_sym_block(fused.dt))
        r = torch.arange(N, device=cuda)
        kept = torch.arange(fused.ld, device=cuda)[None, :] < torch.clamp((r // blk + 1) * blk, max=fused.ld)[:, None]
        assert torch.equal(fused._cov[:, kept], plain._cov[:, kept]), step
    for name in ("mean", "var", "l", "u"):
        assert torch.equal(getattr(fused, name), getattr(plain, name)), name
