"""Test harness: BASELINE config 3 row-sharded over the ranks, driven kernel by kernel through the
engine (talgp.engine.DeviceGP with a communicator).  Used by tests/test_gpu_sharded.py (threads on one
GPU) and tools/mp_parity.py (torchrun, one rank per GPU); bench.py goes through the drop-in API instead."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from talgp.dist import ShardedGP


class ShardedITLBench:
    """BASELINE config 3 row-sharded over the ranks: what bench.py times at --gpus > 1."""

    def __init__(self, domain, A, table, w, world, rank, dev, conditioning="incremental", comm=None,
                 storage="full"):
        from talgp import _abi
        self.w = w
        N = w["N"]
        self.gp = ShardedGP(N, 1, dtype=torch.float64, device=dev, comm=comm, storage=storage)
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
        self.ev, self.pass_ms, self._tctx = [], [], None
        self._y_host = torch.zeros((1,), dtype=torch.float64).pin_memory()
        self._y_dev = torch.zeros((1,), dtype=torch.float64, device=dev)

    def score(self):
        return self.gp.score_itl_directed(self.roi, self.m, incremental=self.inc, roi_key=self)

    def _select(self):
        return self.gp.argmax_candidates(self.score(), None, 1, None)

    def breakdown(self, steps: int = 10):
        """Per-phase device times of one step (CUDA events; diagnostic)."""
        gp = self.gp
        names = ["score", "argmax_local", "combine_records", "exchange_row", "step_vectors",
                 "cond_downdate", "rank1_update"]
        acc = np.zeros(len(names))
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record()
            F = self.score(); ev[1].record()
            rec = gp.masked_argmax(F, None, 1, None, idx_offset=gp.cand0, l_offset=gp.cand0); ev[2].record()
            if gp.comm is not None:
                gp.combine_records(rec)
                rec = gp._combined
            ev[3].record()
            idx_dev = rec[0:1]
            gp.exchange_row(idx_dev); ev[4].record()
            gp.step_vectors(idx_dev, 0, self.table, self.w["N"], True, self.beta); ev[5].record()
            if gp._cond is not None and gp._cond["valid"]:
                gp._cond_downdate(idx_dev, 0)
            ev[6].record()
            gp.rank1_update(); ev[7].record()
            torch.cuda.synchronize()
            acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))])
        return dict(zip(names, (acc / steps * 1e3).round(1).tolist()))  # microseconds

    def step(self, record: bool) -> None:
        """One step, everything on the device.  With a live conditioning state the whole launch
        sequence is one host call (tal_ctx_step); otherwise (first step, re-base) the Python sequence."""
        gp = self.gp
        st = gp._cond
        if self.inc and gp._ctx_usable() and st is not None and st["valid"]:
            gp.ctx_step(self.table, self.w["N"], self.beta, 1, None, None)
            return
        rec = self._select()
        idx_dev = rec[0:1]
        if record and not (self.inc and gp._ctx_usable()):
            gp.exchange_row(idx_dev)
            gp.step_vectors(idx_dev, 0, self.table, self.w["N"], True, self.beta)
            if gp._cond is not None and gp._cond["valid"]:
                gp._cond_downdate(idx_dev, 0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); passed = gp.rank1_update(); e1.record()
            if passed:  # a pass over the matrix was launched (every update_block-th observation)
                self.ev.append((e0, e1))
        else:
            gp.observe(idx_dev, self.table, self.beta, y_gather=True, y_stride=self.w["N"])

    def timing_begin(self) -> None:
        """CUDA events around every pass over the matrix from here on (C-side, csrc/stepctx.cu)."""
        gp = self.gp
        self.pass_ms, self.ev = [], []
        self._tctx = gp._ctx_get() if (self.inc and gp._ctx_usable()) else None
        if self._tctx is not None:
            gp.lib.tal_ctx_timing_begin(C.byref(self._tctx))

    def timing_end(self) -> None:
        gp = self.gp
        if self._tctx is not None:
            from talgp import _abi
            ms, up, n = (C.c_double * 64)(), (C.c_longlong * 64)(), C.c_longlong(0)
            _abi.check(gp.lib.tal_ctx_timing_read(C.byref(self._tctx), ms, up, C.byref(n),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            self.pass_ms = [float(ms[i]) for i in range(int(n.value))]
            self.pass_updates = [int(up[i]) for i in range(int(n.value))]
        else:
            self.pass_ms = [a.elapsed_time(b) for a, b in self.ev]

    def flush(self, record: bool = False) -> None:
        """Apply the pending updates (end of a timed region: no work is left behind)."""
        gp = self.gp
        if getattr(self, "_tctx", None) is not None and gp._pending:
            from talgp import _abi
            self._tctx.pending = gp._pending
            _abi.check(gp.lib.tal_ctx_flush(C.byref(self._tctx), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            gp._ctx_done(self._tctx)
        elif record and gp._pending:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gp.flush(); e1.record()
            self.ev.append((e0, e1))
        else:
            gp.flush()

    def rank1_ms(self) -> float:
        return float(np.mean(self.pass_ms)) if self.pass_ms else 0.0

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
        self.gp.flush()
        t1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1)], device=self.gp.device, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)
