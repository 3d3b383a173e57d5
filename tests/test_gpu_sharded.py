"""The row-sharded path (talgp/dist.py) on ONE GPU: G host threads drive one
shard each through the very code that runs under NCCL (ShardedGP /
ShardedITLBench), with the two exchange steps emulated between the threads.
The result must reproduce the single-GPU engine: covariance rows bit for bit,
scores and selections within tolerance."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run_sharded(cuda, G, domain, A, table, w, steps, conditioning, exchange="thread", storage="full"):
    from talgp.dist import LocalPeerComm, ThreadComm
    from sharded_harness import ShardedITLBench
    if exchange == "peer":
        # the shards' kernels wait on each other while they share ONE GPU here: a cudaFree issued by the
        # caching allocator (device-wide synchronisation) in one thread would stall against a spinning
        # kernel of the other until its time-out, so start from an empty cache (allocations then map
        # fresh memory, which does not synchronise).  Between processes on separate GPUs this cannot occur.
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    shared = {"slots": [None] * G, "barrier": threading.Barrier(G)}
    out = [None] * G
    errs = []

    def body(rank):
        comm = ThreadComm(shared, rank, G) if exchange == "thread" else LocalPeerComm(shared, rank, G)
        b = ShardedITLBench(domain, A, table, w, G, rank, cuda, conditioning, comm=comm, storage=storage)
        sel, Fs = [], []
        for t in range(steps):
            F = b.score()
            Fs.append(F.clone())
            rec = b.gp.argmax_candidates(F, None, 1, None)
            sel.append(int(rec[0].item()))
            b.gp.observe(rec[0:1], b.table, b.beta, y_gather=True, y_stride=w["N"])
        if exchange == "peer":
            torch.cuda.current_stream().synchronize()
            assert comm.error() == 0
        out[rank] = (b, sel, Fs)

    def worker(rank):
        try:
            torch.cuda.set_device(cuda)
            if exchange == "peer":  # the shards' kernels wait on each other: one stream per shard
                with torch.cuda.stream(torch.cuda.Stream(device=cuda)):
                    body(rank)
                    torch.cuda.current_stream().synchronize()
            else:
                body(rank)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            shared["barrier"].abort()

    ths = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    return out


@pytest.mark.parametrize("storage", ["full", "lower"])
@pytest.mark.parametrize("exchange", ["thread", "peer"])
@pytest.mark.parametrize("conditioning", ["incremental", "recompute"])
@pytest.mark.parametrize("G", [2, 3])
def test_sharded_itl_matches_single_gpu(cuda, G, conditioning, exchange, storage):
    """exchange="thread": host-emulated collectives on one stream; "peer": the NVLink
    peer-memory protocol of csrc/peer.cu (mailboxes, P2P stores, release/acquire
    sequence numbers) between shards running concurrently on their own streams."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_problem
    from talgp import _abi
    from talgp.engine import DeviceGP
    # lower-block storage: N spans three 1024-column blocks so that the shards hold pieces of a row
    w = (dict(N=700, m=90, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=6) if storage == "full" else
         dict(N=2500, m=150, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=8))
    domain, A, table = make_problem(w)
    N, steps = w["N"], 8
    # single-GPU engine
    gp = DeviceGP(N, 1, device=cuda, storage=storage)
    dom = torch.as_tensor(domain, device=cuda).contiguous()
    gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, w["lengthscale"], dom)
    gp.rho.fill_(w["rho"]); gp.init_bounds([1.0])
    roi = torch.as_tensor(A, device=cuda)
    tab = torch.as_tensor(table, device=cuda)
    sel1, F1 = [], []
    for t in range(steps):
        F = gp.score_itl_directed(roi, len(A), incremental=(conditioning == "incremental"), roi_key=roi)
        F1.append(F.clone())
        rec = gp.masked_argmax(F, None, 1, None)
        sel1.append(int(rec[0].item()))
        gp.observe(rec[0:1].clone(), tab, [1.0], y_gather=True, y_stride=N)
    out = _run_sharded(cuda, G, domain, A, table, w, steps, conditioning, exchange, storage)
    # every rank saw the same selections, equal to the single-GPU run
    for b, sel, Fs in out:
        assert sel == sel1
    for t in range(steps):
        Fcat = torch.cat([out[r][2][t] for r in range(G)])
        torch.testing.assert_close(Fcat, F1[t], rtol=1e-7, atol=1e-10)
    rows = torch.cat([out[r][0].gp.cov[0] for r in range(G)])
    if storage == "lower":  # only the blocks on or below the diagonal are maintained
        blk = int(gp.lib.tal_sym_block(gp.dt))
        r = torch.arange(N, device=cuda)
        kept = torch.arange(gp.ld, device=cuda)[None, :] < torch.clamp((r // blk + 1) * blk, max=gp.ld)[:, None]
        assert torch.equal(rows[kept], gp.cov[0][kept])
    else:
        assert torch.equal(rows, gp.cov[0])  # the shards' rows ARE the single-GPU matrix
    for r in range(G):
        g = out[r][0].gp
        assert torch.equal(g.mean, gp.mean) and torch.equal(g.var, gp.var)
        assert torch.equal(g.l, gp.l) and torch.equal(g.u, gp.u)


def test_sharded_vtl_columns_match_rows(cuda):
    """VTL over a shard's own candidates through the symmetric column gather."""
    from talgp import _abi
    from talgp.dist import shard_rows
    from talgp.engine import DeviceGP
    import oracle as O
    N, q = 500, 2
    rng = np.random.default_rng(4)
    domain = rng.random((N, 3))
    covs = [O.Matern52(lengthscale=0.3).covariance(domain), O.Gaussian(lengthscale=0.2).covariance(domain)]
    full = DeviceGP(N, q, device=cuda)
    for i in range(q):
        full.load_covariance(i, torch.as_tensor(covs[i]))
    full.rho.fill_(0.1)
    roi = torch.as_tensor(np.sort(rng.choice(N, 70, replace=False)), device=cuda)
    F = full.score_rows(_abi.RULE_VTL, roi, 70)
    parts = []
    for r in range(3):
        row0, n_loc = shard_rows(N, 3, r)
        sh = DeviceGP(N, q, device=cuda, row0=row0, n_loc=n_loc)
        for i in range(q):
            sh.load_covariance(i, torch.as_tensor(covs[i]))
        sh.rho.fill_(0.1)
        parts.append(sh.score_rows(_abi.RULE_VTL, roi, 70, sharded_cols=True))
    torch.testing.assert_close(torch.cat(parts), F, rtol=1e-10, atol=0)


@pytest.mark.parametrize("name,storage", [("c1", "full"), ("c2", "full"), ("c3", "full"), ("c4", "full"),
                                          ("c1", "lower"), ("c2", "lower"), ("c3", "lower"), ("c4", "lower")])
def test_drop_in_api_row_sharded(cuda, name, storage):
    """The `lib.*` classes over a row-sharded model (3 ranks emulated by threads):
    same selections, scores and posterior as the single-GPU model, in both storage
    modes (lower-block storage: the VTL-type rules sum what every shard keeps)."""
    import problems
    from talgp.dist import ThreadComm
    p = problems.ALL[name]()
    T = min(p["T"], 8)
    m1, a1, f1 = problems.build_cuda(p, storage=storage)
    ref = []
    for t in range(T):
        r = a1.step(f1)
        ref.append((m1.last_argmax.idx, r["F"].clone()))
    G = 3
    shared = {"slots": [None] * G, "barrier": threading.Barrier(G)}
    out, errs = [None] * G, []

    def worker(rank):
        try:
            torch.cuda.set_device(cuda)
            m, a, f = problems.build_cuda(p, comm=ThreadComm(shared, rank, G), storage=storage)
            got = []
            for t in range(T):
                r = a.step(f)
                got.append((m.last_argmax.idx, r["F"].clone()))
            out[rank] = (m, got)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            shared["barrier"].abort()

    ths = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    for rank in range(G):
        m, got = out[rank]
        for (i1, F1), (i2, F2) in zip(ref, got):
            assert i1 == i2
            fin = torch.isfinite(F1)
            torch.testing.assert_close(F2[fin], F1[fin], rtol=1e-7, atol=1e-10)
            assert torch.equal(torch.isneginf(F1), torch.isneginf(F2))
        torch.testing.assert_close(m.means, m1.means, rtol=1e-10, atol=1e-13)
        torch.testing.assert_close(m.variances, m1.variances, rtol=1e-10, atol=1e-13)
        row0, n_loc = m.row_range
        if storage == "full":
            torch.testing.assert_close(m.covariances, m1.covariances[:, row0:row0 + n_loc], rtol=1e-10, atol=1e-13)
        else:  # the kept blocks (`m.covariances` is a collective here: see the test below)
            full = m1.covariances[:, row0:row0 + n_loc]
            mine = m._gp.cov[:, :, : m.n]
            low = torch.tril(torch.ones((m.n, m.n), dtype=torch.bool, device=cuda))[row0:row0 + n_loc]
            torch.testing.assert_close(mine[:, low], full[:, low], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("storage", ["lower", "full"])
def test_sharded_model_serves_full_rows_and_sqd_correlation(cuda, storage):
    """`model.covariances` (this rank's rows, complete) and `sqd_correlation` on a row-sharded model: in
    lower-block storage the mirror images of a rank's rows live on the later ranks and are assembled by the
    symmetric row gather + a sum over the shards (collective); same values as the single-GPU model."""
    import problems
    from talgp.dist import ThreadComm
    G = 3
    p = problems.c4(T=4)
    model1, alg1, f1 = problems.build_cuda(p, storage=storage)
    for t in range(4):
        alg1.step(f1)
    cov1 = model1.covariances.clone()
    X = model1.domain[[3, 50, 177, 399]]
    A = model1.domain[[0, 9, 120]]
    sq1 = model1.sqd_correlation(X, A)
    shared = {"slots": [None] * G, "barrier": threading.Barrier(G)}
    out, errs = [None] * G, []

    def worker(rank):
        try:
            torch.cuda.set_device(cuda)
            comm = ThreadComm(shared, rank, G)
            model, alg, f = problems.build_cuda(p, comm=comm, storage=storage)
            for t in range(4):
                alg.step(f)
            cov = model.covariances.clone()   # collective in lower storage
            sq = model.sqd_correlation(X, A)  # collective
            out[rank] = (model._gp.row0, model._gp.n_loc, cov, sq)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            shared["barrier"].abort()

    ths = [threading.Thread(target=worker, args=(r,)) for r in range(G)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    if errs:
        raise errs[0]
    for row0, n_loc, cov, sq in out:
        torch.testing.assert_close(cov, cov1[:, row0: row0 + n_loc], rtol=1e-12, atol=1e-14)
        torch.testing.assert_close(sq, sq1, rtol=1e-10, atol=1e-14)
