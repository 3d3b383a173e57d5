"""Temporally blocked posterior update (csrc/rankb.cu): keeping the (k, w) of up to b observations
pending, correcting the observed row on the fly and applying all of them in one pass must be
BIT-IDENTICAL to a pass over the matrix after every observation (update_block = 1) — every entry
sees the same sequence of rounded operations."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(cuda, N, q, dtype, storage, block):
    from talgp import _abi
    from talgp.engine import DeviceGP
    rng = np.random.default_rng(3)
    dom = torch.as_tensor(rng.random((N, 3)), device=cuda)
    gp = DeviceGP(N, q, dtype=dtype, device=cuda, storage=storage)
    gp.set_update_block(block)
    for i in range(q):
        gp.build_gram(i, _abi.KERNEL_MATERN52 if i % 2 == 0 else _abi.KERNEL_GAUSSIAN, 1.0 + 0.5 * i, 0.3 + 0.1 * i, dom)
    gp.rho.fill_(0.1)
    gp.init_bounds([2.0] * q)
    return gp


@pytest.mark.parametrize("block", [2, 5, 8, 16, 32])
@pytest.mark.parametrize("storage", ["full", "lower"])
@pytest.mark.parametrize("dtype,N", [(torch.float64, 1500), (torch.float32, 2100)])
def test_blocked_update_is_bit_identical(cuda, dtype, N, storage, block):
    q, T = 2, 37 if block <= 16 else 70
    ref = _make(cuda, N, q, dtype, storage, 1)
    gp = _make(cuda, N, q, dtype, storage, block)
    rng = np.random.default_rng(11)
    for t in range(T):
        idx = int(rng.integers(N)) if t % 7 else 5  # the same point observed repeatedly, too
        y = rng.standard_normal(q)
        ref.observe(idx, y, [2.0] * q)
        gp.observe(idx, y, [2.0] * q)
        assert torch.equal(gp.kbuf, ref.kbuf) and torch.equal(gp.wbuf, ref.wbuf), t  # the corrected row
        if t == 20:  # a reader in the middle applies whatever is pending (here fewer than `block`)
            assert torch.equal(gp.covariance_view(0), ref.covariance_view(0))
    assert gp._pending == (T - 21) % block or block == 1
    for name in ("mean", "var", "l", "u"):
        assert torch.equal(getattr(gp, name), getattr(ref, name)), name
    if storage == "lower":
        blk = int(gp.lib.tal_sym_block(gp.dt))
        r = torch.arange(N, device=cuda)
        kept = torch.arange(gp.ld, device=cuda)[None, :] < torch.clamp((r // blk + 1) * blk, max=gp.ld)[:, None]
        assert torch.equal(gp.cov[:, kept], ref.cov[:, kept])
    else:
        assert torch.equal(gp.cov, ref.cov)
    assert gp._pending == 0
    for i in range(q):
        assert torch.equal(torch.diagonal(gp.cov[i, :, :N]).to(torch.float64), gp.var[i])


def test_blocked_update_with_incremental_itl_and_general_m(cuda):
    """Scores of the incremental directed ITL and a general-m conditioning in between are the same
    with and without deferral (the general-m path applies the pending updates first)."""
    N, m = 1300, 150
    ref = _make(cuda, N, 1, torch.float64, "lower", 1)
    gp = _make(cuda, N, 1, torch.float64, "lower", 8)
    rng = np.random.default_rng(2)
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=cuda)
    for t in range(21):
        Fr = ref.score_itl_directed(roi, m, incremental=True, roi_key=roi)
        Fg = gp.score_itl_directed(roi, m, incremental=True, roi_key=roi)
        assert torch.equal(Fr, Fg), t
        idx = int(torch.argmax(Fr))
        y = rng.standard_normal(1)
        ref.observe(idx, y, [2.0])
        gp.observe(idx, y, [2.0])
        if t == 10:
            oi = torch.tensor([3, 700, 1201], device=cuda)
            for g in (ref, gp):
                g.condition(0, oi, torch.tensor([0.1, -0.2, 0.3], device=cuda, dtype=torch.float64),
                            torch.full((3,), 0.1, device=cuda, dtype=torch.float64))
    assert torch.equal(gp.mean, ref.mean) and torch.equal(gp.var, ref.var)
