"""CPU tests of host-side logic that needs no GPU: the row / candidate partitions of the sharded
model in both storage modes, and — as a NumPy statement of what csrc/rankb.cu does — the identity
behind the temporally blocked update: deferring b rank-1 updates, correcting the observed row on the
fly and applying them in one pass gives bit for bit the matrix of b separate passes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))


@pytest.mark.parametrize("N,world", [(65536, 8), (65536, 2), (200000, 8), (2500, 3), (700, 2), (64, 4)])
@pytest.mark.parametrize("storage", ["full", "lower"])
def test_shard_bounds_partition_and_balance(N, world, storage):
    from talgp.dist import owner_of, shard_bounds, shard_rows
    b = shard_bounds(N, world, storage=storage)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == N
    assert all(b[g] <= b[g + 1] for g in range(world))
    assert all(b[g] % 64 == 0 for g in range(world))  # update CTAs (32 rows) and 16-byte pieces never straddle
    for g in range(world):
        assert shard_rows(N, world, g, storage=storage) == (b[g], b[g + 1] - b[g])
    for idx in (0, N // 3, N - 1):
        g = owner_of(N, world, idx, storage=storage)
        assert b[g] <= idx < b[g + 1]
    if N >= 65536:  # balanced update traffic: kept entries per shard within 3 % of even
        ld = -(-N // 32) * 32
        r = np.arange(N)
        per_row = np.minimum((r // 32 + 1) * 32, ld) if storage == "lower" else np.full(N, ld)
        cost = np.array([per_row[b[g]:b[g + 1]].sum() for g in range(world)], dtype=float)
        assert cost.max() / cost.mean() < 1.03, cost / cost.mean()


def _rank1_step(S, x, rho):
    """Sigma' = Sigma - k^T w with w = (k / L) / L, two roundings per entry (rank1.cu / step_vectors)."""
    k = S[x].copy()
    L = np.sqrt(k[x] + rho * rho)
    w = (k / L) / L
    return S - np.outer(k, w), k, w  # np.outer rounds the product, the subtraction rounds again


@pytest.mark.parametrize("b", [2, 5, 16])
def test_deferred_updates_are_bit_identical_numpy(b):
    rng = np.random.default_rng(0)
    n = 60
    X = rng.random((n, 3))
    d = np.sqrt(((X[:, None] - X[None]) ** 2).sum(-1))
    S0 = (1 + np.sqrt(5) * d / 0.4 + 5 / 3 * (d / 0.4) ** 2) * np.exp(-np.sqrt(5) * d / 0.4)
    xs = rng.integers(n, size=37)
    # reference: a pass after every observation
    S = S0.copy()
    ks, ws = [], []
    for x in xs:
        S, k, w = _rank1_step(S, int(x), 0.1)
        ks.append(k); ws.append(w)
    # deferred: base matrix + pending (k, w); the observed row corrected on the fly
    base, pend = S0.copy(), []
    for t, x in enumerate(xs):
        k = base[int(x)].copy()
        for (kp, wp) in pend:  # tal_pending_correct_row
            k = k - kp[int(x)] * wp
        assert np.array_equal(k, ks[t])
        L = np.sqrt(k[int(x)] + 0.1 * 0.1)
        w = (k / L) / L
        assert np.array_equal(w, ws[t])
        pend.append((k, w))
        if len(pend) == b:  # tal_rankb_update
            for (kp, wp) in pend:
                base = base - np.outer(kp, wp)
            pend = []
    for (kp, wp) in pend:
        base = base - np.outer(kp, wp)
    assert np.array_equal(base, S)
