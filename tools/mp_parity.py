"""Parity of the row-sharded path under real multi-process execution (torchrun, one rank per
GPU, CUDA-IPC peer mailboxes or NCCL) against the single-GPU engine, on a small problem:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/mp_parity.py [full|lower] [incremental|recompute]

Every rank also runs the single-GPU engine on its own device and compares selections, scores of
its candidates, the replicated vectors and its covariance rows (the kept blocks, bit for bit)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "transductive-active-learning_b200")):
    sys.path.insert(0, p)
from bench import make_problem  # noqa: E402
from talgp import _abi  # noqa: E402
from talgp.dist import default_comm  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sharded_harness import ShardedITLBench  # noqa: E402
from talgp.engine import DeviceGP  # noqa: E402


def main():
    storage = sys.argv[1] if len(sys.argv) > 1 else "lower"
    conditioning = sys.argv[2] if len(sys.argv) > 2 else "incremental"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    w = dict(N=3300, m=200, d=4, lengthscale=0.25, rho=0.1, beta=1.0, grid=8)
    domain, A, table = make_problem(w)
    N, steps = w["N"], 10
    comm = default_comm()
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
    # single-GPU engine on this rank's device
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
    for name in ("mean", "var", "l", "u"):
        assert torch.equal(getattr(g, name), getattr(gp, name)), name
    mine, ref = g.cov[0], gp.cov[0, g.row0: g.row0 + g.n_loc]
    if storage == "lower":
        blk = int(gp.lib.tal_sym_block(gp.dt))
        r = torch.arange(g.row0, g.row0 + g.n_loc, device=dev)
        kept = torch.arange(gp.ld, device=dev)[None, :] < torch.clamp((r // blk + 1) * blk, max=gp.ld)[:, None]
        assert torch.equal(mine[kept], ref[kept])
    else:
        assert torch.equal(mine, ref)
    dist.barrier()
    if rank == 0:
        print(f"mp parity ok: world={world} storage={storage} conditioning={conditioning} exchange="
              f"{'peer' if getattr(comm, 'peer', False) else 'nccl'} selections={sel}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
