"""Parity of the row-sharded path under real multi-process execution (torchrun, one rank per
GPU, CUDA-IPC peer mailboxes or NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port 29533 tools/mp_parity.py small [full|lower] [incremental|recompute]
    ... tools/mp_parity.py fused [full|lower] [q]        # the fused step (tal_ctx_step), q GPs
    ... tools/mp_parity.py c5 [steps] | c3 [steps]       # FULL SIZE against the low-rank replay oracle

small / fused: every rank also runs the single-GPU engine on its own device and compares selections,
scores of its candidates, the replicated vectors and its covariance rows (the kept blocks, bit for bit).
c5 / c3: BASELINE configs[4] / [2] through the drop-in API on G ranks; rank 0 replays the selections
with oracle/replay.py and checks means / variances / bounds of all N candidates at 1e-10 and the
scores + selections at a few steps."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transductive-active-learning_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from bench import WORKLOADS, make_problem  # noqa: E402
from talgp import _abi  # noqa: E402
from talgp.dist import default_comm  # noqa: E402
from talgp.engine import DeviceGP  # noqa: E402
from sharded_harness import ShardedITLBench  # noqa: E402


def _kept(gp, dev):
    blk = int(gp.lib.tal_sym_block(gp.dt))
    r = torch.arange(gp.row0, gp.row0 + gp.n_loc, device=dev)
    return torch.arange(gp.ld, device=dev)[None, :] < torch.clamp((r // blk + 1) * blk, max=gp.ld)[:, None]


def _compare(g, gp, storage, dev):
    for name in ("mean", "var", "l", "u"):
        assert torch.equal(getattr(g, name), getattr(gp, name)), name
    mine, ref = g.cov, gp.cov[:, g.row0: g.row0 + g.n_loc]
    if storage == "lower":
        k = _kept(g, dev)
        assert torch.equal(mine[:, k], ref[:, k])
    else:
        assert torch.equal(mine, ref)


def small(storage, conditioning, rank, world, dev, comm):
    w = dict(N=3300, m=200, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=8)
    domain, A, table = make_problem(w)
    N, steps = w["N"], 10
    b = ShardedITLBench(domain, A, table, w, world, rank, dev, conditioning, comm=comm, storage=storage)
    sel, Fs = [], []
    for t in range(steps):
        F = b.score()
        Fs.append(F.clone())
        rec = b.gp.argmax_candidates(F, None, 1, None)
        sel.append(int(rec[0].item()))
        b.gp.observe(rec[0:1], b.table, b.beta, y_gather=True, y_stride=N)
    torch.cuda.synchronize()
    if getattr(comm, "peer", False):
        assert comm.error() == 0
    gp = DeviceGP(N, 1, device=dev, storage=storage)
    gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, w["lengthscale"], torch.as_tensor(domain, device=dev).contiguous())
    gp.rho.fill_(w["rho"]); gp.init_bounds([1.0])
    roi, tab = torch.as_tensor(A, device=dev), torch.as_tensor(table, device=dev)
    sel1 = []
    g = b.gp
    for t in range(steps):
        F = gp.score_itl_directed(roi, len(A), incremental=(conditioning == "incremental"), roi_key=roi)
        torch.testing.assert_close(Fs[t], F[g.cand0: g.cand0 + g.n_cand], rtol=1e-7, atol=1e-10)
        rec = gp.masked_argmax(F, None, 1, None)
        sel1.append(int(rec[0].item()))
        gp.observe(rec[0:1].clone(), tab, [1.0], y_gather=True, y_stride=N)
    assert sel == sel1, (rank, sel, sel1)
    _compare(g, gp, storage, dev)
    return f"storage={storage} conditioning={conditioning} selections={sel}"


def fused(storage, q, rank, world, dev, comm):
    """tal_ctx_step on every rank (records published by fused_finish, combined in the push kernel)."""
    N, m, T = 4100, 220, 40
    rng = np.random.default_rng(6)
    domain = torch.as_tensor(rng.random((N, 3)), device=dev)
    roi = torch.as_tensor(np.sort(rng.choice(N, m, replace=False)), device=dev)
    table = torch.as_tensor(rng.standard_normal((q, N)), device=dev)

    def make(c):
        gp = DeviceGP(N, q, device=dev, storage=storage, arith="fma", comm=c)
        for i in range(q):
            gp.build_gram(i, _abi.KERNEL_MATERN52 if i % 2 == 0 else _abi.KERNEL_GAUSSIAN, 1.0 + 0.5 * i,
                          0.3 + 0.1 * i, domain)
        gp.rho.fill_(0.1)
        if q > 1:
            gp.mean[1:].fill_(4.0)  # constraint GPs start pessimistically safe
        gp.init_bounds([2.0] * q)
        return gp

    def run(gp):
        sel = []
        for t in range(T):
            st = gp._cond
            if st is None or not st["valid"]:
                F = gp.score_itl_directed(roi, m, incremental=True, roi_key=roi)
                rec = gp.argmax_candidates(F, None, 1, None)
                sel.append(int(rec[0].item()))
                gp.observe(rec[0:1].clone(), table, [2.0] * q, y_gather=True, y_stride=N)
            else:
                if t % 5 == 0:  # (reading it first takes the stand-alone combine path; otherwise the push combines)
                    sel.append(int(gp.ctx_select(1, None, None)[0].item()))
                gp.ctx_step(table, N, [2.0] * q, 1, None, None)
        gp.flush()
        torch.cuda.synchronize()
        return sel
    g = make(comm)
    sel = run(g)
    if getattr(comm, "peer", False):
        assert comm.error() == 0
    one = make(None)
    sel1 = run(one)
    assert sel == sel1, (rank, sel, sel1)
    _compare(g, one, storage, dev)
    return f"fused storage={storage} q={q} selections={sel[:8]}..."


def full_size(name, steps, rank, world, dev, comm):
    import argparse
    import oracle as O
    from oracle.replay import LowRankReplay
    from bench import build_itl
    from lib.function import TableFunction
    w = WORKLOADS[name]
    args = argparse.Namespace(storage="lower", dtype="f64", arith="fma", conditioning="incremental", workload=name)
    domain, A, table, model, alg = build_itl(w, args, dev, comm)
    alg.gather_F = True
    f = TableFunction(model.domain, table)
    N = w["N"]
    check = {0, 1, 2, 15, 16, 17, steps - 1}
    sel, Fs = [], {}
    t0 = time.time()
    for t in range(steps):
        out = alg.step(f)
        sel.append(int(model.last_argmax.idx))
        if t in check and rank == 0:
            Fs[t] = out["F"].cpu().numpy()
    model._gp.flush()
    torch.cuda.synchronize()
    gpu_s = time.time() - t0
    if getattr(comm, "peer", False):
        assert comm.error() == 0
    gp = model._gp
    if rank == 0:
        rep = LowRankReplay(O.Matern52(lengthscale=w["lengthscale"]), domain, w["rho"], w["beta"])
        rng = np.random.default_rng(5)
        worst = 0.0
        for t in range(steps):
            if t in Fs:
                F = Fs[t]
                top = np.argsort(-F)[:48]
                C = np.unique(np.concatenate([rng.choice(N, 1024, replace=False), top, [sel[t]]]))
                Fo = rep.itl_directed(A, C)
                np.testing.assert_allclose(F[C], Fo, rtol=1e-6, atol=1e-9, err_msg=f"step {t}")
                worst = max(worst, float(np.max(np.abs(F[C] - Fo) / np.maximum(np.abs(Fo), 1e-3))))
                mx, s_ = float(Fo.max()), float(Fo[np.searchsorted(C, sel[t])])
                assert s_ >= mx - 1e-9 * max(1.0, abs(mx)), (t, sel[t], s_, mx)
            rep.observe(sel[t], float(table[0, sel[t]]))
        np.testing.assert_allclose(gp.mean[0].cpu().numpy(), rep.mean, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(gp.var[0].cpu().numpy(), rep.var, rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(gp.l[0].cpu().numpy(), rep.l, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(gp.u[0].cpu().numpy(), rep.u, rtol=1e-10, atol=1e-12)
        msg = (f"full size {name}: N={N} m={w['m']} steps={steps} ({gpu_s:.1f} s on the GPUs); means / variances / l / u "
               f"of all candidates within 1e-10 of the replay oracle; scores at steps {sorted(Fs)} within 1e-6 "
               f"(worst {worst:.1e}); every checked selection maximises the oracle objective; "
               f"{len(set(sel))} distinct points")
    else:
        msg = ""
    # the replicated state is identical on all ranks
    digest = torch.stack([gp.mean.sum(), gp.var.sum(), gp.l.sum(), gp.u.sum()])
    dig_all = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(dig_all, digest)
    assert all(bool(torch.equal(d_, dig_all[0])) for d_ in dig_all)
    return msg


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "small"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = default_comm()
    if mode in ("full", "lower"):  # (round-1 command line: storage first)
        msg = small(mode, sys.argv[2] if len(sys.argv) > 2 else "incremental", rank, world, dev, comm)
    elif mode == "small":
        msg = small(sys.argv[2] if len(sys.argv) > 2 else "lower", sys.argv[3] if len(sys.argv) > 3 else "incremental",
                    rank, world, dev, comm)
    elif mode == "fused":
        msg = fused(sys.argv[2] if len(sys.argv) > 2 else "lower", int(sys.argv[3]) if len(sys.argv) > 3 else 2,
                    rank, world, dev, comm)
    else:
        msg = full_size(mode, int(sys.argv[2]) if len(sys.argv) > 2 else 40, rank, world, dev, comm)
    dist.barrier()
    if rank == 0:
        print(f"mp parity ok: world={world} exchange={'peer' if getattr(comm, 'peer', False) else 'nccl'} {msg}",
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
