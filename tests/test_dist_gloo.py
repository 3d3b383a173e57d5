"""world_size-2 CPU tests (gloo) of the host-side logic of the row-sharded path:
row partition, the sum-as-broadcast of the observed covariance row, and the
all-gather + lexicographic combine of the per-shard arg-max records.  The
per-shard arithmetic is done by the NumPy oracle here (the CUDA kernels need a
GPU); the exchange steps are the ones talgp/dist.py runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from talgp.dist import combine_records, owner_of, shard_rows, status_of, sum_broadcast
        N = 200
        rng = np.random.default_rng(0)  # same problem on every rank
        domain = rng.random((N, 2))
        cov = O.Matern52(lengthscale=0.4).covariance(domain)
        row0, n_loc = shard_rows(N, world, rank)
        assert (row0, n_loc) == ((0, 128), (128, 72))[rank]
        local = cov[row0:row0 + n_loc].copy()  # this rank's row block
        full = O.MarginalModel(0, domain, [O.GaussianDistribution(np.zeros(N), cov.copy())], beta=np.array([2.0]),
                               noise_rate=O.HomoscedasticNoise(1, np.array([0.1])), use_objective_as_constraint=True)
        mean, var = np.zeros(N), np.diag(cov).copy()
        l, u = mean - 2 * np.sqrt(var), mean + 2 * np.sqrt(var)
        l[:] = 0.5  # make everything safe so that the arg-max is defined
        full._l[:] = 0.5
        for t in range(6):
            # ---- per-shard scoring + arg-max record, all-gather, combine ----
            F = O.itl_undirected(full) + np.round(rng.standard_normal(N), 1)  # ties across shards
            F_loc = F[row0:row0 + n_loc]
            safe_loc = l[row0:row0 + n_loc] >= 0
            sf = np.where(safe_loc, F_loc, -np.inf)
            mx = np.nanmax(sf)
            ties = np.nonzero(sf == mx)[0]
            rec = torch.tensor([float(ties[0] + row0), mx, float(len(ties)), 7.0], dtype=torch.float64)
            recs = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(recs, rec)
            idx, val, cnt, flags = combine_records([(int(r[0]), float(r[1]), int(r[2]), int(r[3])) for r in recs])
            tie_set, gmx = full.objective_tie_set(F)
            assert idx == tie_set[0] and val == gmx and cnt == len(tie_set) and status_of(flags) == 0
            # ---- the observed row: only the owner fills the buffer, the sum broadcasts it ----
            k = torch.zeros(N, dtype=torch.float64)
            if owner_of(N, world, idx) == rank:
                k[:] = torch.from_numpy(local[idx - row0])
            sum_broadcast(k)
            np.testing.assert_allclose(k.numpy(), full.covariances[0][idx], rtol=1e-12, atol=1e-15)
            if t == 0:
                assert np.array_equal(k.numpy(), full.covariances[0][idx])  # x + 0 = x bit for bit
            # ---- every shard updates its own rows; the O(N) vectors are replicated ----
            kk = k.numpy()
            y = float(rng.standard_normal())
            s = kk[idx] + 0.1 ** 2
            w = (kk / np.sqrt(s)) / np.sqrt(s)
            local -= np.outer(kk[row0:row0 + n_loc], w)
            full.step_idx(np.array([idx]), np.array([[y]]), np.array([[0.1]]))
            np.testing.assert_allclose(local, full.covariances[0][row0:row0 + n_loc], rtol=1e-12, atol=1e-15)
        # empty shards combine correctly
        assert combine_records([(-1, -np.inf, 0, 0), (5, 1.0, 2, 7)]) == (5, 1.0, 2, 7)
        assert status_of(0) == 2 and status_of(1) == 3 and status_of(3) == 4
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_rows_cover_the_domain():
    sys.path.insert(0, os.path.join(ROOT, "transductive-active-learning_b200"))
    from talgp.dist import owner_of, shard_rows
    for N in (1, 63, 64, 200, 65536, 200000):
        for world in (1, 2, 4, 8):
            blocks = [shard_rows(N, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == N
            for (a, n), (b, _) in zip(blocks, blocks[1:]):
                assert a + n == b or n == 0 or b == N
            for idx in (0, N // 2, N - 1):
                r = owner_of(N, world, idx)
                assert blocks[r][0] <= idx < blocks[r][0] + blocks[r][1]
