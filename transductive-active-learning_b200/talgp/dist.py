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


# ------------------------------------------------------------------ device side
class ShardedGP:
    """DeviceGP restricted to this rank's row block + the two exchange steps."""

    def __init__(self, N: int, q: int, dtype=torch.float64, device=None, comm=None,
                 rank1_variant: int = 0):
        from .engine import DeviceGP
        self.comm = comm if comm is not None else NcclComm()
        self.world, self.rank = self.comm.world, self.comm.rank
        self.row0, self.n_loc = shard_rows(N, self.world, self.rank)
        self.gp = DeviceGP(N, q, dtype=dtype, device=device, row0=self.row0, n_loc=self.n_loc,
                           rank1_variant=rank1_variant)
        self.N, self.q = N, q
        dev = self.gp.device
        self._recs = torch.zeros((self.world, 4), dtype=torch.int64, device=dev)
        self._combined = torch.zeros((4,), dtype=torch.int64, device=dev)
        self._cond = None

    # -- construction ---------------------------------------------------------
    def build_gram(self, i: int, kind: int, variance: float, lengthscale: float,
                   domain: torch.Tensor) -> None:
        gp = self.gp
        gp.var[i].zero_()
        gp.build_gram(i, kind, variance, lengthscale, domain)
        self.comm.sum_(gp.var[i])  # each rank contributed its own diagonal block

    # -- arg-max ----------------------------------------------------------------
    def masked_argmax(self, F_loc: torch.Tensor, first_constraint: int,
                      sample: Optional[torch.Tensor]) -> torch.Tensor:
        """Local reduce -> all-gather of the 32-byte records -> combine.  Returns
        the combined device record (int64[4]; [0] is the global index)."""
        from . import _abi
        from .engine import _ptr, _stream
        gp = self.gp
        smp = sample[self.row0:self.row0 + self.n_loc] if sample is not None else None
        rec = gp.masked_argmax(F_loc, None, first_constraint, smp, idx_offset=self.row0,
                               l_offset=self.row0)
        if self.world == 1:
            return rec
        self.comm.gather_records(rec, self._recs)
        _abi.check(gp.lib.tal_argmax_combine(_ptr(self._recs), self.world, _ptr(self._combined),
                                             _stream()), "tal_argmax_combine")
        return self._combined

    # -- update -------------------------------------------------------------------
    def observe(self, idx_dev: torch.Tensor, y_dev: torch.Tensor, beta: Sequence[float],
                y_gather: bool, y_stride: int) -> None:
        gp = self.gp
        gp.extract_row(idx_dev)
        self.comm.sum_(gp.kbuf)
        gp.step_vectors(idx_dev, 0, y_dev, y_stride, y_gather, beta)
        if self._cond is not None:
            self._cond_downdate(idx_dev)
        gp.rank1_update()

    # -- directed ITL over this shard's candidates ------------------------------------
    def score_itl_directed(self, roi_idx: torch.Tensor, m: int, jitter: float = 1e-5,
                           incremental: bool = True) -> torch.Tensor:
        from . import _abi
        from .engine import _ptr, _stream, round_up
        gp, lib = self.gp, self.gp.lib
        dev = gp.device
        F = torch.empty((self.n_loc,), dtype=torch.float64, device=dev)
        st = self._cond
        if st is None or not incremental:
            m_pad, ncols = round_up(m, 128), round_up(max(self.n_loc, 1), 64)
            if st is None:
                st = dict(V=torch.empty((self.q, m_pad, ncols), dtype=torch.float64, device=dev),
                          cs=torch.empty((self.q, ncols), dtype=torch.float64, device=dev),
                          pcol=torch.empty((m_pad,), dtype=torch.float64, device=dev),
                          coef=torch.empty((int(lib.tal_cond_coef_bytes(m_pad)) // 8,),
                                           dtype=torch.float64, device=dev),
                          m_pad=m_pad, ncols=ncols)
            S = gp._scratch("cv_S", (m_pad, m_pad))
            work = gp._scratch("cv_work", (int(lib.tal_potrf_work_bytes(m_pad)) // 8,))
            n_split = int(max(1, min(64, m_pad // 64, -(-1184 // max(1, ncols // 512)))))
            part = gp._scratch("cv_part", (n_split, ncols))
            for i in range(self.q):
                V = st["V"][i]
                _abi.check(lib.tal_gather_roi_sharded(
                    _ptr(gp.cov[i]), gp.ld, self.row0, self.n_loc, _ptr(roi_idx), m, m_pad,
                    float(jitter), _ptr(S), m_pad, _ptr(V), ncols, ncols, gp.dt, _stream()),
                    "tal_gather_roi_sharded")
                self.comm.sum_(S)  # Sigma_AA assembled from the owners' rows
                _abi.check(lib.tal_potrf_lower(_ptr(S), m_pad, m_pad, _ptr(work), _stream()),
                           "tal_potrf_lower")  # replicated: identical on every rank
                _abi.check(lib.tal_trsm_lower(_ptr(S), m_pad, m_pad, _ptr(work), _ptr(V), ncols,
                                              ncols, _stream()), "tal_trsm_lower")
                _abi.check(lib.tal_colsumsq_dense(_ptr(V), ncols, m_pad, ncols, _ptr(part), n_split,
                                                  _ptr(st["cs"][i]), _stream()), "tal_colsumsq_dense")
            self._cond = st if incremental else None
        sl = slice(self.row0, self.row0 + self.n_loc)
        for i in range(self.q):
            _abi.check(lib.tal_score_itl_directed(
                _ptr(gp.var[i, sl.start:]), _ptr(st["cs"][i]), _ptr(gp.rho[i, sl.start:]), self.n_loc,
                1 if i > 0 else 0, _ptr(F), _stream()), "tal_score_itl_directed")
        return F

    def _cond_downdate(self, idx_dev: torch.Tensor) -> None:
        from . import _abi
        from .engine import _ptr, _stream
        gp, st = self.gp, self._cond
        for i in range(self.q):
            V = st["V"][i]
            _abi.check(gp.lib.tal_cond_extract_col(
                _ptr(V), st["ncols"], st["m_pad"], self.row0, self.n_loc, _ptr(idx_dev), 0,
                _ptr(st["pcol"]), _stream()), "tal_cond_extract_col")
            self.comm.sum_(st["pcol"])
            _abi.check(gp.lib.tal_cond_update(
                _ptr(V), st["ncols"], st["m_pad"], st["ncols"], self.n_loc, self.row0,
                _ptr(st["pcol"]), _ptr(gp.kbuf[i]), _ptr(gp.wbuf[i]), _ptr(gp.rho[i]),
                _ptr(idx_dev), 0, _ptr(st["coef"]), _ptr(st["cs"][i]), gp.dt, _stream()),
                "tal_cond_update")

    def score_rows(self, rule: int, roi_idx: torch.Tensor, m: int, **kw) -> torch.Tensor:
        """VTL / CTL / MM-ITL / TruVar over this shard's candidates (no communication)."""
        return self.gp.score_rows(rule, roi_idx, m, sharded_cols=True, **kw)


class ShardedITLBench:
    """BASELINE config 3 row-sharded over the ranks: what bench.py times at --gpus > 1."""

    def __init__(self, domain, A, table, w, world, rank, dev, conditioning="incremental", comm=None):
        from . import _abi
        self.w = w
        N = w["N"]
        self.sg = ShardedGP(N, 1, dtype=torch.float64, device=dev, comm=comm)
        self.gp = self.sg.gp
        dom = torch.as_tensor(domain, device=dev).contiguous()
        self.sg.build_gram(0, _abi.KERNEL_MATERN52, 1.0, w["lengthscale"], dom)
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

    def _select(self):
        F = self.sg.score_itl_directed(self.roi, self.m, incremental=self.inc)
        return self.sg.masked_argmax(F, 1, None)

    def step(self, record: bool) -> None:
        rec = self._select()
        idx_dev = rec[0:1]
        gp = self.gp
        gp.extract_row(idx_dev)
        self.sg.comm.sum_(gp.kbuf)
        gp.step_vectors(idx_dev, 0, self.table, self.w["N"], True, self.beta)
        if self.sg._cond is not None:
            self.sg._cond_downdate(idx_dev)
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
            self.sg.observe(rec[0:1], self._y_dev, self.beta, False, 1)
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=self.gp.device, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
