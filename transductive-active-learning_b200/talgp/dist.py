"""Row-sharded discrete GP over the GPUs of one box: one process per GPU,
`torch.distributed` (NCCL over NVLink / NVSwitch) for the two exchange steps.

Partition: GPU g holds rows [row0_g, row0_g + n_loc_g) of every
N x N posterior covariance; the O(qN) vectors are replicated.  Per BO step:

  * scoring is local to each shard's own candidates (symmetric column gather);
  * arg-max: every rank reduces its candidates, the 32-byte records are
    all-gathered and combined with the same lexicographic rule on every rank;
  * the observed row k = Sigma[x*, :] is completed on every rank by a sum
    all-reduce of a buffer that only the owner filled (no host-known root, so
    the step needs no host synchronisation), together with V[:, x*] when the
    target-set conditioning is kept incrementally;
  * the rank-1 update then streams each shard's rows independently.

The host-side helpers (`shard_rows`, `combine_records`, `sum_broadcast`) work on
CPU tensors with the gloo backend as well; tests/test_dist_gloo.py runs them at
world_size 2 without a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ------------------------------------------------------------------ host logic
def shard_rows(N: int, world: int, rank: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous row block of `rank`: (row0, n_loc); blocks are multiples of
    `align` rows except the last one."""
    per = -(-N // world)
    per = -(-per // align) * align
    row0 = min(N, rank * per)
    return row0, max(0, min(N, row0 + per) - row0)


def owner_of(N: int, world: int, idx: int, align: int = 64) -> int:
    per = -(-N // world)
    per = -(-per // align) * align
    return int(idx) // per


def sum_broadcast(buf: torch.Tensor, group=None) -> torch.Tensor:
    """Broadcast by summation: exactly one rank holds non-zeros (x + 0 = x bit
    for bit), so no rank needs to know the root on the host."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def combine_records(recs: Sequence[Tuple[int, float, int, int]]) -> Tuple[int, float, int, int]:
    """Python statement of tal_argmax_combine: records (idx, val, n_ties, flags);
    max val, then min idx, tie counts add, flags OR."""
    idx, val, cnt, flags = -1, -np.inf, 0, 0
    for (i, v, c, f) in recs:
        flags |= f
        if c == 0:
            continue
        if cnt == 0 or v > val:
            idx, val, cnt = i, v, c
        elif v == val:
            idx, cnt = min(idx, i), cnt + c
    return idx, val, cnt, flags


def status_of(flags: int) -> int:
    if not flags & 1:
        return 2
    if not flags & 2:
        return 3
    if not flags & 4:
        return 4
    return 0


class NcclComm:
    """The two exchange steps over torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def sum_(self, buf: torch.Tensor) -> torch.Tensor:
        return sum_broadcast(buf, self.group)

    def gather_records(self, rec: torch.Tensor, out: torch.Tensor) -> None:
        dist.all_gather_into_tensor(out.view(-1), rec, group=self.group)

    def all_gather_(self, part: torch.Tensor, out: torch.Tensor) -> None:
        dist.all_gather_into_tensor(out, part, group=self.group)


class ThreadComm:
    """The same exchange steps between `world` Python threads that drive one
    shard each on ONE GPU (all on the same stream, so kernels serialise in issue
    order and never wait on one another).  Used to test the sharded path on a
    single GPU: the ShardedGP code is the one that runs under NCCL."""

    def __init__(self, shared: dict, rank: int, world: int):
        self.shared, self.rank, self.world = shared, rank, world

    def sum_(self, buf: torch.Tensor) -> torch.Tensor:
        sh = self.shared
        sh["slots"][self.rank] = buf
        sh["barrier"].wait()
        tot = sh["slots"][0].clone()
        for r in range(1, self.world):
            tot += sh["slots"][r]
        sh["barrier"].wait()  # every thread has issued its reads
        buf.copy_(tot)
        sh["barrier"].wait()
        return buf

    def gather_records(self, rec: torch.Tensor, out: torch.Tensor) -> None:
        sh = self.shared
        sh["slots"][self.rank] = rec
        sh["barrier"].wait()
        for r in range(self.world):
            out[r].copy_(sh["slots"][r])
        sh["barrier"].wait()

    def all_gather_(self, part: torch.Tensor, out: torch.Tensor) -> None:
        sh = self.shared
        sh["slots"][self.rank] = part
        sh["barrier"].wait()
        n = part.numel()
        for r in range(self.world):
            out[r * n:(r + 1) * n].copy_(sh["slots"][r])
        sh["barrier"].wait()


# ------------------------------------------------------------------ device side
def ShardedGP(N: int, q: int, dtype=torch.float64, device=None, comm=None, rank1_variant: int = 0):
    """A DeviceGP holding this rank's row block (talgp.engine.DeviceGP with a communicator)."""
    from .engine import DeviceGP
    return DeviceGP(N, q, dtype=dtype, device=device, rank1_variant=rank1_variant,
                    comm=comm if comm is not None else NcclComm())


class ShardedITLBench:
    """BASELINE config 3 row-sharded over the ranks: what bench.py times at --gpus > 1."""

    def __init__(self, domain, A, table, w, world, rank, dev, conditioning="incremental", comm=None):
        from . import _abi
        self.w = w
        N = w["N"]
        self.gp = ShardedGP(N, 1, dtype=torch.float64, device=dev, comm=comm)
        dom = torch.as_tensor(domain, device=dev).contiguous()
        self.gp.build_gram(0, _abi.KERNEL_MATERN52, 1.0, w["lengthscale"], dom)
        self.gp.rho.fill_(w["rho"])
        self.beta = [w["beta"]]
        self.gp.init_bounds(self.beta)
        self.roi = torch.as_tensor(A, device=dev)
        self.m = int(len(A))
        self.table_host = table
        self.table = torch.as_tensor(table, device=dev)
        self.inc = conditioning == "incremental"
        self.ev = []
        self._y_host = torch.zeros((1,), dtype=torch.float64).pin_memory()
        self._y_dev = torch.zeros((1,), dtype=torch.float64, device=dev)

    def score(self):
        return self.gp.score_itl_directed(self.roi, self.m, incremental=self.inc, roi_key=self)

    def _select(self):
        return self.gp.argmax_candidates(self.score(), None, 1, None)

    def breakdown(self, steps: int = 10):
        """Per-phase device times of one step (CUDA events; diagnostic)."""
        gp = self.gp
        names = ["score", "argmax_local", "gather+combine", "extract_row", "sum(kbuf)", "step_vectors",
                 "cond_extract", "sum(pcol)", "cond_update", "rank1_update"]
        acc = np.zeros(len(names))
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            F = self.score(); ev[1].record()
            sl = slice(gp.row0, gp.row0 + gp.n_loc)
            rec = gp.masked_argmax(F, None, 1, None, idx_offset=gp.row0, l_offset=gp.row0); ev[2].record()
            if gp.comm is not None:
                from . import _abi
                from .engine import _ptr, _stream
                gp.comm.gather_records(rec, gp._recs)
                _abi.check(gp.lib.tal_argmax_combine(_ptr(gp._recs), gp.comm.world, _ptr(gp._combined), _stream()))
                rec = gp._combined
            ev[3].record()
            idx_dev = rec[0:1]
            gp.extract_row(idx_dev); ev[4].record()
            gp._sum(gp.kbuf); ev[5].record()
            gp.step_vectors(idx_dev, 0, self.table, self.w["N"], True, self.beta); ev[6].record()
            st = gp._cond
            from . import _abi
            from .engine import _ptr, _stream
            _abi.check(gp.lib.tal_cond_extract_col(_ptr(st["V"][0]), st["ncols"], st["m_pad"], gp.row0, gp.n_loc,
                                                   _ptr(idx_dev), 0, _ptr(st["pcol"]), _stream())); ev[7].record()
            gp._sum(st["pcol"]); ev[8].record()
            _abi.check(gp.lib.tal_cond_update(_ptr(st["V"][0]), st["ncols"], st["m_pad"], st["ncols"], gp.n_loc,
                                              gp.row0, _ptr(st["pcol"]), _ptr(gp.kbuf[0]), _ptr(gp.wbuf[0]),
                                              _ptr(gp.rho[0]), _ptr(idx_dev), 0, _ptr(st["coef"]),
                                              _ptr(st["cs"][0]), int(gp.cond_chunks), _ptr(st["chunk"]),
                                              gp.dt, _stream())); ev[9].record()
            gp.rank1_update(); ev[10].record()
            torch.cuda.synchronize()
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
        return dict(zip(names, (acc / steps * 1e3).round(1).tolist()))  # microseconds

    def step(self, record: bool) -> None:
        rec = self._select()
        idx_dev = rec[0:1]
        gp = self.gp
        gp.extract_row(idx_dev)
        gp._sum(gp.kbuf)
        gp.step_vectors(idx_dev, 0, self.table, self.w["N"], True, self.beta)
        if gp._cond is not None and gp._cond["valid"]:
            gp._cond_downdate(idx_dev, 0)
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gp.rank1_update(); e1.record()
            self.ev.append((e0, e1))
        else:
            gp.rank1_update()

    def rank1_ms(self) -> float:
        return float(np.mean([a.elapsed_time(b) for a, b in self.ev])) if self.ev else 0.0

    def e2e(self, steps: int) -> float:
        """The same steps with a host black box: every rank reads the combined
        record back (D2H), looks the observation up on the host and copies it in
        from pinned memory (H2D)."""
        torch.cuda.synchronize()
        dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            rec = self._select()
            r = self.gp.read_result(rec)
            assert r.status == 0
            self._y_host[0] = float(self.table_host[0, r.idx])
            self._y_dev.copy_(self._y_host, non_blocking=True)
            self.gp.observe(rec[0:1], self._y_dev, self.beta)
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=self.gp.device, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
