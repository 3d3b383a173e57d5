"""Device-resident state of q Gaussian processes over a finite domain, and the
kernels that act on it (through the C ABI of include/tal_b200.h).

This is the engine underneath the drop-in `lib.model` / `lib.algorithms`
classes.  State layout in HBM (DESIGN.md "Data layout"):

  cov   [q, n_loc, ld]  storage dtype (fp64 or fp32), row-major, ld = N rounded
                        up to 32 elements, pad columns zero; rows
                        [row0, row0 + n_loc) of every GP's N x N posterior
                        covariance (one GPU: row0 = 0, n_loc = N)
  mean, var, l, u, rho  [q, N] fp64, replicated on every shard
  kbuf, wbuf            [q, ld] storage dtype: the observed row k and the
                        weights w = (k / L) / L of the current step

Reference correspondence: MarginalModel's `_means`, `_covariances`, `_l`,
`_u` (lib/model/marginal.py:43-51); `var` replaces the strided
`vmap(jnp.diag)` gather of `variances` (:103-106) and is kept bit-identical
to the matrix diagonal by the update kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi

_DT = {torch.float64: _abi.F64, torch.float32: _abi.F32}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def round_up(x: int, k: int) -> int:
    return (x + k - 1) // k * k


def nearest_index(domain: torch.Tensor, A: torch.Tensor) -> torch.Tensor:
    """get_indices(domain, A) (lib/utils.py:53-56): argmin_j ||domain[j] - A[k]||_2,
    first minimum wins.  Returns int64 [m] on the device."""
    _abi.require_device()
    lib = _abi.load()
    domain = domain.to(dtype=torch.float64).contiguous()
    A = A.to(device=domain.device, dtype=torch.float64).contiguous()
    m = A.shape[0]
    out = torch.empty((m,), dtype=torch.int64, device=domain.device)
    _abi.check(lib.tal_nearest_index(_ptr(domain), domain.shape[0], _ptr(A), m, domain.shape[1],
                                     _ptr(out), _stream()), "tal_nearest_index")
    return out


class DeviceGP:
    def __init__(self, N: int, q: int, dtype: torch.dtype = torch.float64,
                 device: Optional[torch.device] = None, row0: int = 0,
                 n_loc: Optional[int] = None, rank1_variant: int = 0, comm=None,
                 storage: str = "full", arith: Optional[str] = None):
        """comm: None (one GPU) or a communicator of talgp.dist (NcclComm /
        ThreadComm): the covariance is then row-sharded over comm.world ranks and
        this object holds rank comm.rank's block.
        storage: "full" keeps every entry of the N x N matrices like the reference;
        "lower" maintains only the blocks on or below the diagonal (the matrix is
        symmetric), which halves the HBM traffic of every posterior update; readers
        of full rows are served through the mirrored entries (include/tal_b200.h,
        "lower-block storage").
        arith: "two_rounding" (default; c - rn(k w), bit-compatible with NumPy / an unfused
        dot + subtract) or "fma" (one fused multiply-subtract per entry and update: half the
        fp64 issue slots of the deferred pass); env TAL_UPDATE_ARITH.  The matrix, the pending
        rows and `var` always share the mode, so var == diag(cov) holds bit for bit in both."""
        _abi.require_device()
        if storage not in ("full", "lower"):
            raise ValueError("storage must be 'full' or 'lower'")
        self.storage, self.lower = storage, storage == "lower"
        self._upper_stale = False
        self.comm = comm if (comm is not None and comm.world > 1) else None
        from .dist import shard_bounds
        if self.comm is not None:
            self.bounds = shard_bounds(int(N), comm.world, storage=storage)
            row0, n_loc = self.bounds[comm.rank], self.bounds[comm.rank + 1] - self.bounds[comm.rank]
        else:
            self.bounds = [0, int(N)]
        self.lib = _abi.load()
        if q < 1 or q > _abi.TAL_MAX_Q:
            raise ValueError(f"q must be in [1, {_abi.TAL_MAX_Q}]")
        if dtype not in _DT:
            raise ValueError("dtype must be torch.float64 or torch.float32")
        self.N, self.q = int(N), int(q)
        self.dtype, self.dt = dtype, _DT[dtype]
        arith = arith or os.environ.get("TAL_UPDATE_ARITH", "two_rounding")
        if arith not in ("two_rounding", "fma"):
            raise ValueError("arith must be 'two_rounding' or 'fma'")
        self.arith = arith
        self.dt_upd = self.dt | (_abi.ARITH_FMA if arith == "fma" else 0)  # dtype argument of the update calls
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda", torch.cuda.current_device())
        self.row0 = int(row0)
        self.n_loc = self.N if n_loc is None else int(n_loc)
        # candidates this rank scores (and holds conditioning state for): its own rows, except
        # in lower-block storage, where the rows are split by AREA and the candidates evenly
        self.cand_bounds = self.bounds
        if self.comm is not None and self.lower:
            self.cand_bounds = shard_bounds(self.N, comm.world, storage="full")
        if self.comm is not None:
            self.cand0 = self.cand_bounds[comm.rank]
            self.n_cand = self.cand_bounds[comm.rank + 1] - self.cand0
        else:
            self.cand0, self.n_cand = self.row0, self.n_loc
        self.ld = round_up(self.N, 32)
        self.rank1_variant = rank1_variant
        dev = self.device
        self._cov = torch.zeros((q, self.n_loc, self.ld), dtype=dtype, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        self.mean = torch.zeros((q, N), **f64)
        self.var = torch.zeros((q, N), **f64)
        self.l = torch.zeros((q, N), **f64)
        self.u = torch.zeros((q, N), **f64)
        self.rho = torch.zeros((q, N), **f64)
        self.kbuf = torch.zeros((q, self.ld), dtype=dtype, device=dev)
        self.wbuf = torch.zeros((q, self.ld), dtype=dtype, device=dev)
        # temporally blocked posterior update (csrc/rankb.cu): the (k, w) of up to `update_block`
        # observations are kept pending and applied to the matrix in one pass
        self.update_block = max(1, min(_abi.TAL_MAX_PENDING, int(os.environ.get("TAL_UPDATE_BLOCK", "16"))))
        self._pending = 0
        if self.update_block > 1:
            self._alloc_pending()
        self.scal = torch.zeros((q,), **f64)
        self._argmax_scratch = torch.zeros((int(self.lib.tal_argmax_scratch_bytes()),),
                                           dtype=torch.uint8, device=dev)
        self._result = torch.zeros((4,), dtype=torch.int64, device=dev)  # tal_argmax_result
        self._result_host = torch.zeros((4,), dtype=torch.int64).pin_memory()
        self._y_host = torch.zeros((q,), dtype=torch.float64).pin_memory()
        self._y_dev = torch.zeros((64, q), **f64)  # ring of device slots for host-supplied observations
        self._y_slot, self._y_event = 0, None
        self._ws = {}  # named scratch tensors, grown on demand
        self._ws_views = {}  # name -> (tensor, shape, dtype, view) of the last request
        self._cond = None  # incremental conditioning state (score_itl_directed)
        self._ctx = None   # tal_step_ctx of the fused step (built on demand)
        self._ctx_st = None
        self._gen = 0      # bumped whenever a buffer the context points to may have been replaced
        self._ctx_gen = -1
        self.fused = os.environ.get("TAL_FUSED", "1") != "0"  # False: the kernel-by-kernel Python sequence
        self.cond_chunks = 0  # tal_cond_update: 0 = choose, 1 = single pass, > 1 = row-parallel
        # incremental conditioning: "append" (read-only pass + two appended rows per observation,
        # csrc/condapp.cu) or "downdate" (rewrites V every step, csrc/condinc.cu)
        self.cond_form = os.environ.get("TAL_COND_FORM", "append")
        self.cond_capacity = int(os.environ.get("TAL_COND_CAPACITY", "1024"))  # observations before a rebase
        self._peer = None  # NVLink peer-memory exchange state (talgp.dist.PeerViews)
        if self.comm is not None:
            self._recs = torch.zeros((self.comm.world, 4), dtype=torch.int64, device=dev)
            self._combined = torch.zeros((4,), dtype=torch.int64, device=dev)
            if getattr(self.comm, "peer", False):
                self._peer_attach(0)

    @property
    def sharded(self) -> bool:
        return self.comm is not None

    @property
    def cov(self) -> torch.Tensor:
        """[q, n_loc, ld] covariance rows with every pending update applied."""
        self.flush()
        return self._cov

    def set_update_block(self, b: int) -> None:
        """Number of observations whose covariance updates are applied in one pass (1 = a pass per
        observation, like the reference)."""
        self.flush()
        self._gen += 1
        self.update_block = max(1, min(_abi.TAL_MAX_PENDING, int(b)))
        if self.update_block > 1:
            self._alloc_pending()

    def _alloc_pending(self) -> None:
        self._gen = getattr(self, "_gen", 0) + 1
        self._alloc_pending_impl()

    def _alloc_pending_impl(self) -> None:
        """Pending (k, w) slots (the pass is compiled for 2, 4, 8 or 16 of them; slots beyond the
        number in use are not read)."""
        slots = 2
        while slots < self.update_block:
            slots *= 2
        self._Kp = torch.zeros((slots, self.q, self.ld), dtype=self.dtype, device=self.device)
        self._Wp = torch.zeros((slots, self.q, self.ld), dtype=self.dtype, device=self.device)

    def flush(self) -> bool:
        """Apply the pending rank-1 updates to the matrix in one pass (bit-identical to applying
        them one by one); returns whether a pass was launched."""
        p = self._pending
        if p == 0:
            return False
        _abi.check(self.lib.tal_rankb_update(
            _ptr(self._cov), self.cov_stride_q, self.ld, self.row0, self.n_loc, self.N, self.q,
            _ptr(self._Kp), _ptr(self._Wp), self.q * self.ld, p, int(self.lower), self.dt_upd, _stream()),
            "tal_rankb_update")
        self._pending = 0  # (slots >= p are never read: nothing to clear)
        if self.lower:
            self._upper_stale = True
        return True

    def _peer_attach(self, m_pad: int) -> None:
        """(Re)allocate the peer mailboxes; the observed-row buffer kbuf and the
        conditioning column then live inside this rank's mailbox, where the
        owner of the row stores them (csrc/peer.cu).  Collective."""
        self._sel_discard()
        self._gen += 1
        self._ctx = None  # (its cached mailbox pointers are about to be freed)
        self._peer = self.comm.attach(self.q, self.ld, self.dtype, int(m_pad), self.device)
        self.kbuf = self._peer.kbuf
        self._bounds_c = (C.c_longlong * len(self.bounds))(*self.bounds)

    def exchange_row(self, idx_dev: Optional[torch.Tensor], idx_host: int = 0) -> None:
        """kbuf <- cov[:, idx, :] on every rank (the row lives on one shard) and
        scal <- mean[:, idx]; in peer mode also pcol <- V[:, :, idx] when the
        incremental conditioning is live."""
        if self._peer is None:
            self.extract_row(idx_dev, idx_host)  # (one GPU: corrected for the pending updates there)
            if self.comm is not None:
                self._sum(self.kbuf)
                self.correct_row(idx_dev, idx_host)
            return
        st = self._cond if (self._cond is not None and self._cond["valid"]) else None
        pv, comm = self._peer, self.comm
        _abi.check(self.lib.tal_peer_push_row(
            comm.table_ptr, comm.world, comm.rank, self._bounds_c, _ptr(self._cov), self.cov_stride_q,
            self.ld, self.N, self.q, int(self.lower), _ptr(idx_dev), int(idx_host), pv.kbuf_off,
            _ptr(st["V"]) if st is not None else None,
            st["cap"] * st["ncols"] if st is not None else 0, st["ncols"] if st is not None else 0,
            st["rows"] if st is not None else 0, pv.pcol_off, pv.pcol.shape[1], self.cand0, self.n_cand,
            self.dt, _stream()),
            "tal_peer_push_row")
        _abi.check(self.lib.tal_peer_wait_row(comm.box_ptr, comm.world, _ptr(idx_dev), int(idx_host), _ptr(self.mean),
                                              self.N, self.q, _ptr(self.scal), _stream()), "tal_peer_wait_row")
        self.correct_row(idx_dev, idx_host)

    # ------------------------------------------------- one step as one host call
    def _ctx_usable(self) -> bool:
        """The C-side step sequence (csrc/stepctx.cu + fused.cu) covers one GPU and the peer-memory
        sharded model, with no conditioning state or the append form."""
        if not self.fused or (self.comm is not None and self._peer is None):
            return False
        st = self._cond
        return st is None or st["append"]

    def _ctx_get(self) -> "_abi.StepCtx":
        st = self._cond
        c = self._ctx
        if c is not None and self._ctx_gen == self._gen and self._ctx_st is st:
            # nothing the context points to has been replaced since it was built (every site that (re)allocates one
            # of those buffers bumps `_gen`): skip the pointer-by-pointer key
            c.pending = self._pending
            c.rows = st["rows"] if st is not None else 0
            c.cond_valid = int(st is not None and st["valid"])
            return c
        F = self._scratch("ctx_F", (max(self.n_cand, 1),))
        key = (None if st is None else (st["V"].data_ptr(), st["cap"], st["cs"].data_ptr(), st["fz"].data_ptr()),
               None if self._peer is None else (self._peer.kbuf.data_ptr(), self.comm._table.data_ptr()),
               self.update_block, self._Kp.data_ptr() if self.update_block > 1 else 0, self._cov.data_ptr(),
               F.data_ptr(), self.arith)
        c = getattr(self, "_ctx", None)
        if c is None or self._ctx_key != key:
            if c is not None:
                self._sel_discard()  # (a published selection belongs to the old context)
            old = c
            c = _abi.StepCtx()
            if old is not None:  # (library-owned side stream / events of the overlap: keep them)
                c.side_stream, c.ev_fork, c.ev_join = old.side_stream, old.ev_fork, old.ev_join
            # observe / scoring-pass overlap on a side stream -- not in the single-GPU emulation of the peer protocol,
            # where creating a stream while another shard's kernel spins on this one stalls until its time-out
            # (the hazard of tests/conftest.py; between processes on separate GPUs nothing is shared)
            # OPT-IN (TAL_OVERLAP=1): measured +3 % at 4 GPUs, but one full-size parity run at 4 ranks failed with it
            # (scores of one column tile off by 2e-5 at step 15) -- the signature of the kdiag race fixed since
            # (csrc/peer.cuh); off until that case has been re-run green with the fix
            c.overlap = int(os.environ.get("TAL_OVERLAP", "0") == "1"
                            and not getattr(self.comm, "single_context", False))
            c.cov, c.cov_stride_q, c.ld, c.row0, c.n_loc = self._cov.data_ptr(), self.cov_stride_q, self.ld, self.row0, self.n_loc
            c.N, c.q, c.dtype, c.lower = self.N, self.q, self.dt, int(self.lower)
            c.arith = int(self.arith == "fma")
            for name in ("mean", "var", "l", "u", "rho", "scal", "kbuf", "wbuf"):
                setattr(c, name, getattr(self, name).data_ptr())
            if self.update_block > 1:
                c.Kp, c.Wp = self._Kp.data_ptr(), self._Wp.data_ptr()
                c.pend_stride, c.pend_slots = self.q * self.ld, self._Kp.shape[0]
            c.update_block = self.update_block
            c.cand0, c.n_cand = self.cand0, self.n_cand
            if st is not None:
                c.Wb, c.ldw, c.cap_rows, c.m_pad, c.ncols = st["V"].data_ptr(), st["ncols"], st["cap"], st["m_pad"], st["ncols"]
                c.n_split, c.cs, c.app_scratch, c.pcol = st["n_split"], st["cs"].data_ptr(), st["app"].data_ptr(), st["pcol"].data_ptr()
                c.fz_scratch = st["fz"].data_ptr()
            c.F = F.data_ptr()
            c.argmax_scratch = self._argmax_scratch.data_ptr()
            c.rec, c.combined = self._result.data_ptr(), self._result.data_ptr()
            c.G, c.rank = 1, 0
            if self._peer is not None:
                comm, pv = self.comm, self._peer
                c.G, c.rank = comm.world, comm.rank
                c.mailboxes, c.mailbox = comm._table.data_ptr(), comm._box
                for g, b in enumerate(self.bounds):
                    c.bounds[g] = b
                for g in range(len(self.bounds), 17):
                    c.bounds[g] = self.bounds[-1]
                c.kbuf_off, c.pcol_off = pv.kbuf_off, pv.pcol_off
                c.pcol_peer, c.pcol_stride = pv.pcol.data_ptr(), pv.pcol.shape[1]
                c.combined = self._combined.data_ptr()
            self._ctx, self._ctx_key = c, key
        self._ctx_gen, self._ctx_st = self._gen, st
        c.pending = self._pending
        c.rows = st["rows"] if st is not None else 0
        c.cond_valid = int(st is not None and st["valid"])
        return c

    def _ctx_done(self, c) -> None:
        self._pending = int(c.pending)
        if self._cond is not None:
            self._cond["rows"] = int(c.rows)
            if not c.cond_valid:
                self._cond["valid"] = False
        if self.lower and c.passes:
            self._upper_stale = True

    def _sel_discard(self) -> None:
        """The model is about to change outside the fused step: a selection that the previous
        step made ahead of time is stale.  Sharded: its record was already published to the
        peers, so the exchange is consumed (every rank does the same)."""
        c = getattr(self, "_ctx", None)
        if c is None or c.sel_state == 0:
            return
        if c.sel_state == 1:
            _abi.check(self.lib.tal_peer_combine_records(self.comm.box_ptr, self.comm.world, _ptr(self._combined),
                                                         _stream()), "tal_peer_combine_records")
        c.sel_state, c.scal_ok = 0, 0

    def _ctx_masks(self, c, first_constraint: int, safe, sample) -> None:
        c.safe = safe.data_ptr() if safe is not None else None
        c.sample = sample.data_ptr() if sample is not None else None
        c.first_constraint = int(first_constraint)
        self._ctx_mask_refs = (safe, sample)  # keep the tensors alive while the context points at them

    def ctx_observe(self, idx_dev, idx_host, y_dev, y_stride: int, y_gather: bool,
                    beta: Sequence[float], auto_select: bool = False) -> None:
        """`observe` through tal_ctx_observe: the whole per-observation launch sequence in one call.
        y_dev None: the value is supplied later with `ctx_observe_y`."""
        c = self._ctx_get()
        for i in range(self.q):
            c.beta[i] = float(beta[i])
        c.auto_select = int(bool(auto_select))
        rc = self.lib.tal_ctx_observe(C.byref(c), _ptr(idx_dev), int(idx_host), _ptr(y_dev), int(y_stride),
                                      int(bool(y_gather)), _stream())
        if rc != _abi.TAL_CTX_REBASE:  # (REBASE: applied without the conditioning update, state marked stale)
            _abi.check(rc, "tal_ctx_observe")
        self._ctx_done(c)

    def ctx_observe_prepare(self, beta: Sequence[float]):
        """First half of a value-less `ctx_observe` (see ITL._step): everything the host can do BEFORE it knows
        the selection -- the context with its pointers, beta, the stream handle."""
        c = self._ctx_get()
        for i in range(self.q):
            c.beta[i] = float(beta[i])
        c.auto_select = 0
        return c, _stream()

    def ctx_observe_fire(self, c, stream, idx_ptr: int) -> None:
        """Second half: enqueue the y-independent part of the observation at the index stored at device
        address idx_ptr (the arg-max record): one ctypes call, three launches."""
        rc = self.lib.tal_ctx_observe(C.byref(c), C.c_void_p(idx_ptr), 0, None, 1, 0, stream)
        if rc != _abi.TAL_CTX_REBASE:
            _abi.check(rc, "tal_ctx_observe")
        self._ctx_done(c)

    def ctx_observe_y(self, idx_dev, idx_host, y_dev: torch.Tensor, y_stride: int, y_gather: bool) -> None:
        c = self._ctx
        _abi.check(self.lib.tal_ctx_observe_y(C.byref(c), _ptr(idx_dev), int(idx_host), _ptr(y_dev), int(y_stride),
                                              int(bool(y_gather)), _stream()), "tal_ctx_observe_y")

    def ctx_select(self, first_constraint: int, safe: Optional[torch.Tensor] = None,
                   sample: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Directed-ITL selection over the resident conditioning state (nothing is launched when the
        previous fused step already made it); returns the device record."""
        st = self._cond
        if not (self._ctx_usable() and st is not None and st["valid"]):
            raise _abi.TalError("ctx_select needs a valid resident conditioning state (score_itl_directed(incremental=True))")
        c = self._ctx_get()
        self._ctx_masks(c, first_constraint, safe, sample)
        _abi.check(self.lib.tal_ctx_select(C.byref(c), _stream()), "tal_ctx_select")
        return self.argmax_record()

    def ctx_step(self, y_table: torch.Tensor, y_stride: int, beta: Sequence[float], first_constraint: int,
                 safe: Optional[torch.Tensor] = None, sample: Optional[torch.Tensor] = None) -> bool:
        """One greedy step of directed ITL over the resident conditioning state in one host call
        (select -> exchange -> update -> selection of the next candidate); returns whether a pass
        over the matrix was launched.  The selection is left in `argmax_record()`.  When the
        conditioning factor is full the observation is applied without it and the state marked
        stale: the caller's next scoring pass re-conditions from scratch."""
        st = self._cond
        if not (self._ctx_usable() and st is not None and st["valid"]):
            raise _abi.TalError("ctx_step needs a valid resident conditioning state (score_itl_directed(incremental=True))")
        c = self._ctx_get()
        self._ctx_masks(c, first_constraint, safe, sample)
        c.y_table, c.y_stride = y_table.data_ptr(), int(y_stride)
        for i in range(self.q):
            c.beta[i] = float(beta[i])
        c.auto_select = 1
        passes = int(c.passes)
        rc = self.lib.tal_ctx_step(C.byref(c), _stream())
        if rc != _abi.TAL_CTX_REBASE:
            _abi.check(rc, "tal_ctx_step")
        self._ctx_done(c)
        return int(c.passes) > passes

    def argmax_record(self) -> torch.Tensor:
        return self._combined if self.comm is not None else self._result

    def correct_row(self, idx_dev: Optional[torch.Tensor], idx_host: int = 0) -> None:
        """kbuf holds row idx of the matrix WITHOUT the pending updates: apply them to the row
        (the same rounded operations the deferred pass will apply to those entries)."""
        if self._pending == 0:
            return
        _abi.check(self.lib.tal_pending_correct_row(
            _ptr(self.kbuf), self.ld, self.N, self.q, _ptr(self._Kp), _ptr(self._Wp), self.q * self.ld,
            self._pending, _ptr(idx_dev), int(idx_host), int(self.lower), self.dt_upd, _stream()),
            "tal_pending_correct_row")

    def _sum(self, buf: torch.Tensor) -> None:
        """Complete a buffer that only its owner filled (sum = broadcast)."""
        if self.comm is not None:
            self.comm.sum_(buf)

    def gather_candidates(self, F_loc: torch.Tensor) -> torch.Tensor:
        """[n_cand] per-shard candidate vector -> [N] on every rank."""
        if self.comm is None:
            return F_loc
        b, G = self.cand_bounds, self.comm.world
        per = max(b[g + 1] - b[g] for g in range(G))
        pad = torch.zeros((per,), dtype=F_loc.dtype, device=self.device)
        pad[: self.n_cand] = F_loc
        out = torch.empty((G * per,), dtype=F_loc.dtype, device=self.device)
        self.comm.all_gather_(pad, out)
        if all(b[g + 1] - b[g] == per for g in range(G - 1)):
            return out[: self.N]
        return torch.cat([out[g * per: g * per + b[g + 1] - b[g]] for g in range(G)])

    # ------------------------------------------------------------------ util
    def _scratch(self, name: str, shape, dtype=torch.float64) -> torch.Tensor:
        t = self._ws.get(name)
        hit = self._ws_views.get(name)  # (the same request as last time: the per-step callers)
        if hit is not None and hit[0] is t and hit[1] == shape and hit[2] is dtype:
            return hit[3]
        n = 1
        for d_ in shape:
            n *= int(d_)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty((n,), dtype=dtype, device=self.device)
            self._ws[name] = t
        v = t[:n].view(*shape)
        self._ws_views[name] = (t, tuple(shape), dtype, v)
        return v

    @property
    def cov_stride_q(self) -> int:
        return self.n_loc * self.ld

    def covariance_view(self, i: int) -> torch.Tensor:
        """[n_loc, N] view of GP i's covariance rows (no copy)."""
        self.ensure_full()
        return self._cov[i, :, : self.N]

    def covariances_view(self) -> torch.Tensor:
        """[q, n_loc, N] view of all covariance rows (no copy)."""
        self.ensure_full()
        return self._cov[:, :, : self.N]

    def ensure_full(self) -> None:
        """Lower-block storage: make the blocks above the diagonal valid again
        (mirror of the kept blocks) before a reader of full rows runs."""
        self.flush()
        if not (self.lower and self._upper_stale):
            return
        if self.comm is not None:
            # row-sharded: the mirror images of this rank's rows live in the later rows of OTHER ranks.  Every
            # global row is assembled 256 at a time by the symmetric row gather + a sum over the shards, its owner
            # keeps it.  COLLECTIVE (every rank must call it) and O(N^2) traffic through the all-reduce: a one-off
            # reader (`model.covariances`), not a per-step path.
            for i in range(self.q):
                for a in range(0, self.N, 256):
                    idx = torch.arange(a, min(self.N, a + 256), device=self.device)
                    rows = self.gather_full_rows(i, idx)
                    lo, hi = max(a, self.row0), min(a + idx.numel(), self.row0 + self.n_loc)
                    if lo < hi:
                        self._cov[i, lo - self.row0: hi - self.row0, : self.N].copy_(
                            rows[lo - a: hi - a, : self.N].to(self.dtype))  # (pad columns stay zero)
            self._upper_stale = False
            return
        _abi.check(self.lib.tal_sym_mirror(_ptr(self._cov), self.cov_stride_q, self.ld, self.N, self.q,
                                           self.dt, _stream()), "tal_sym_mirror")
        self._upper_stale = False

    def gather_full_rows(self, i: int, idx: torch.Tensor) -> torch.Tensor:
        """Full rows cov_i[idx, :] ([len(idx), ld] fp64) on every rank, whatever the storage mode and the sharding:
        every shard contributes the pieces it holds (tal_gather_rows[_lower]), a sum over the shards completes
        them.  Collective when row-sharded.  Up to 256 rows per call."""
        self.flush()
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        m = int(idx.numel())
        rows = self._scratch("full_rows", (max(m, 1), self.ld))
        gather = self.lib.tal_gather_rows_lower if self.lower else self.lib.tal_gather_rows
        _abi.check(gather(_ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.N, _ptr(idx), m, _ptr(rows),
                          self.ld, self.dt, _stream()), "tal_gather_rows")
        self._sum(rows)
        return rows[:m]

    def update_bytes(self) -> float:
        """Algorithmic HBM bytes of one posterior update of this shard (read + write
        of every maintained covariance entry): 2 q n_loc ld s in full storage,
        2 q s sum_rows cend(row) in lower-block storage."""
        s = 8 if self.dtype == torch.float64 else 4
        if not self.lower:
            return 2.0 * self.q * self.n_loc * self.ld * s
        blk = int(self.lib.tal_sym_block(self.dt))
        r = np.arange(self.row0, self.row0 + self.n_loc, dtype=np.int64)
        cend = np.minimum((r // blk + 1) * blk, self.ld)
        return 2.0 * self.q * s * float(cend.sum())

    def launch_count(self) -> int:
        return int(self.lib.tal_launch_count())

    # ------------------------------------------------------------ construction
    def build_gram(self, i: int, kind: int, variance: float, lengthscale: float,
                   domain: torch.Tensor) -> None:
        """cov[i] = K(domain, domain)[row0:row0+n_loc] and var[i] = its diagonal.
        (Kernel.covariance, lib/gp/kernels/base.py:24-25.)"""
        assert domain.dtype == torch.float64 and domain.is_contiguous()
        n, d = domain.shape
        assert n == self.N
        self.flush()
        if self.lower:  # only the kept blocks are built; the mirrored ones on demand (tal_sym_mirror)
            self._upper_stale = True
            _abi.check(self.lib.tal_gram_stationary_lower(
                kind, _ptr(domain), n, d, float(variance), float(lengthscale),
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.dt, _stream()),
                "tal_gram_stationary_lower")
        else:
            _abi.check(self.lib.tal_gram_stationary(
                kind, _ptr(domain), n, _ptr(domain), n, d, float(variance), float(lengthscale),
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.dt, _stream()),
                "tal_gram_stationary")
        # diagonal of a stationary kernel: k(x, x), evaluated by the same kernel
        # on a one-row problem per point is wasteful; the diagonal entries of
        # the shard are copied instead and completed by the caller when sharded.
        if self.n_loc == self.N:
            self.var[i].copy_(torch.diagonal(self._cov[i, :, : self.N]).to(torch.float64))
        else:
            sl = slice(self.row0, self.row0 + self.n_loc)
            self.var[i].zero_()
            self.var[i, sl].copy_(
                torch.diagonal(self._cov[i, :, self.row0:self.row0 + self.n_loc]).to(torch.float64))
            self._sum(self.var[i])  # every shard contributed its own diagonal block

    def load_covariance(self, i: int, cov) -> None:
        """Upload a full N x N covariance (host or device) into shard i."""
        src = torch.as_tensor(cov)
        self.flush()
        if self.q == 1:
            self._upper_stale = False  # (with several GPs the flag keeps covering the other ones)
        rows = src[self.row0:self.row0 + self.n_loc]
        self._cov[i, :, : self.N].copy_(rows.to(device=self.device, dtype=self.dtype))
        self.var[i].copy_(torch.diagonal(src).to(device=self.device, dtype=torch.float64))
        if self.dtype != torch.float64:  # keep var == diag(cov) in the storage precision
            self.var[i].copy_(self.var[i].to(self.dtype).to(torch.float64))

    def init_bounds(self, beta: Sequence[float]) -> None:
        """l = mean - beta*stddev, u = mean + beta*stddev (marginal.py:160-164)."""
        b = torch.as_tensor(np.asarray(beta, dtype=np.float64), device=self.device)[:, None]
        sd = torch.sqrt(self.var)
        self.l.copy_(self.mean - b * sd)
        self.u.copy_(self.mean + b * sd)

    # ------------------------------------------------------------------ update
    def extract_row(self, idx_dev: Optional[torch.Tensor], idx_host: int = 0, correct: bool = True) -> None:
        """kbuf <- this shard's part of row idx (zeros elsewhere) and scal <- mean[:, idx].  On one GPU
        the pending updates are applied to the row as well (`correct`); a sharded caller does that
        after the sum over the shards (`exchange_row`)."""
        fn = self.lib.tal_extract_row_lower if self.lower else self.lib.tal_extract_row
        _abi.check(fn(
            _ptr(self._cov), self.cov_stride_q, self.ld, self.row0, self.n_loc, self.N, self.q,
            _ptr(idx_dev), int(idx_host), _ptr(self.mean), _ptr(self.kbuf), _ptr(self.scal),
            self.dt, _stream()), "tal_extract_row")
        if correct and self.comm is None:
            self.correct_row(idx_dev, idx_host)

    def step_vectors(self, idx_dev: Optional[torch.Tensor], idx_host: int, y_dev: torch.Tensor,
                     y_stride: int, y_gather: bool, beta: Sequence[float]) -> None:
        b = (C.c_double * self.q)(*[float(x) for x in beta])
        _abi.check(self.lib.tal_step_vectors(
            _ptr(self.kbuf), _ptr(self.wbuf), self.N, self.q, _ptr(idx_dev), int(idx_host),
            _ptr(y_dev), int(y_stride), int(bool(y_gather)), _ptr(self.rho), b, _ptr(self.scal),
            _ptr(self.mean), _ptr(self.var), _ptr(self.l), _ptr(self.u), self.dt_upd, _stream()),
            "tal_step_vectors")

    def rank1_update(self, variant: Optional[int] = None) -> bool:
        """The covariance update of the current observation (kbuf, wbuf).  With
        update_block > 1 it is queued and applied together with the next ones
        (`flush`); returns whether a pass over the matrix was launched."""
        if self.update_block > 1 and variant is None:
            self._Kp[self._pending].copy_(self.kbuf)
            self._Wp[self._pending].copy_(self.wbuf)
            self._pending += 1
            return self.flush() if self._pending >= self.update_block else False
        self.flush()
        self._rank1_now(variant)
        return True

    def _rank1_now(self, variant: Optional[int] = None) -> None:
        if self.lower:
            self._upper_stale = True
            _abi.check(self.lib.tal_rank1_update_lower(
                _ptr(self._cov), self.cov_stride_q, self.ld, self.row0, self.n_loc, self.N, self.q,
                _ptr(self.kbuf), _ptr(self.wbuf), self.dt_upd, _stream()), "tal_rank1_update_lower")
            return
        _abi.check(self.lib.tal_rank1_update(
            _ptr(self._cov), self.cov_stride_q, self.ld, self.row0, self.n_loc, self.N, self.q,
            _ptr(self.kbuf), _ptr(self.wbuf), self.dt_upd,
            self.rank1_variant if variant is None else variant, _stream()), "tal_rank1_update")

    def stage_y(self, y) -> torch.Tensor:
        """Device copy of an observation y[q]: a device tensor is used as it is, a host value goes through the
        pinned staging buffer into a fresh slot of a device ring (asynchronous H2D, q * 8 bytes)."""
        if isinstance(y, torch.Tensor) and y.is_cuda:
            return y
        if self._y_event is not None:
            self._y_event.synchronize()  # the previous async copy out of the pinned buffer is done
        self._y_host.copy_(torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1)))
        self._y_slot = (self._y_slot + 1) % self._y_dev.shape[0]
        y_dev = self._y_dev[self._y_slot]  # a fresh device slot: earlier steps may still read theirs
        y_dev.copy_(self._y_host, non_blocking=True)
        if self._y_event is None:
            self._y_event = torch.cuda.Event()
        self._y_event.record()
        return y_dev

    def observe(self, idx, y, beta: Sequence[float], y_gather: bool = False,
                y_stride: int = 1, allreduce=None, auto_select: bool = False) -> None:
        """One observation at domain index `idx` (python int or 1-element int64
        device tensor) with values y[q] (host sequence, or device tensor; with
        y_gather a device table [q, y_stride] indexed at idx).
        `allreduce(t)` completes kbuf across shards (sum; only the owner's is non-zero)."""
        if isinstance(idx, torch.Tensor):
            idx_dev, idx_host = idx, 0
        else:
            idx_dev, idx_host = None, int(idx)
        y_dev = self.stage_y(y)
        if allreduce is None and self._ctx_usable():
            self.ctx_observe(idx_dev, idx_host, y_dev, y_stride, y_gather, beta, auto_select=auto_select)
            return
        self._sel_discard()
        if allreduce is not None:
            self.extract_row(idx_dev, idx_host, correct=False)
            allreduce(self.kbuf)
            self.correct_row(idx_dev, idx_host)
        else:
            self.exchange_row(idx_dev, idx_host)
        self.step_vectors(idx_dev, idx_host, y_dev, y_stride, y_gather, beta)
        if self._cond is not None and self._cond["valid"]:
            self._cond_downdate(idx_dev, idx_host)
        self.rank1_update()

    def condition(self, i: int, obs_idx: torch.Tensor, y: torch.Tensor,
                  noise_std: Optional[torch.Tensor], jitter: float = 1e-5,
                  allreduce=None) -> None:
        """GP i <- posterior given m observations (GaussianDistribution.posterior,
        lib/gp/gaussian_distribution.py:100-121), in chunks of <= 64 observations
        (conditioning chunk after chunk is the same posterior)."""
        m_total = int(obs_idx.numel())
        if m_total == 0:  # :23-24 short-circuit
            return
        self._sel_discard()
        self.flush()
        if self._cond is not None:
            self._cond["valid"] = False  # general-m update: re-condition from scratch next time
        obs_idx = obs_idx.to(device=self.device, dtype=torch.int64).contiguous()
        y = y.to(device=self.device, dtype=torch.float64).contiguous()
        if noise_std is not None:
            noise_std = noise_std.to(device=self.device, dtype=torch.float64).contiguous()
        ldb = self.ld
        small = self._scratch("cond_small", (int(self.lib.tal_small_scratch_bytes()) // 8,))
        for s in range(0, m_total, 64):
            e = min(m_total, s + 64)
            m = e - s
            B = self._scratch("cond_B", (m, ldb))
            W = self._scratch("cond_W", (m, ldb))
            oi = obs_idx[s:e]
            gather = self.lib.tal_gather_rows_lower if self.lower else self.lib.tal_gather_rows
            _abi.check(gather(
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.N, _ptr(oi), m, _ptr(B),
                ldb, self.dt, _stream()), "tal_gather_rows")
            if allreduce is not None:
                allreduce(B)
            else:
                self._sum(B)
            _abi.check(self.lib.tal_small_posterior_weights(
                _ptr(B), ldb, self.N, _ptr(oi), m,
                _ptr(noise_std[s:e]) if noise_std is not None else None, float(jitter),
                _ptr(y[s:e]), _ptr(W), _ptr(self.mean[i]), _ptr(self.var[i]), _ptr(small), _stream()),
                "tal_small_posterior_weights")
            update = self.lib.tal_rankm_update_lower if self.lower else self.lib.tal_rankm_update
            _abi.check(update(
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.N, _ptr(B), _ptr(W), ldb,
                m, self.dt, _stream()), "tal_rankm_update")
            if self.lower:
                self._upper_stale = True
        if self.dtype != torch.float64:
            self.var[i].copy_(self.var[i].to(self.dtype).to(torch.float64))

    # ------------------------------------------------------------------- masks
    def bounds_masks(self, first_constraint: int, op_safe: Optional[torch.Tensor] = None,
                     want_pm: bool = True, want_pe: bool = True):
        """(pess, opt, pm, pe, max_l) — marginal.py:171-224.  Masks are uint8 [N]."""
        N = self.N
        u8 = dict(dtype=torch.uint8, device=self.device)
        pess = torch.empty((N,), **u8)
        opt = torch.empty((N,), **u8)
        pm = torch.empty((N,), **u8) if want_pm else None
        pe = torch.empty((N,), **u8) if want_pe else None
        max_l = torch.empty((1,), dtype=torch.float64, device=self.device)
        _abi.check(self.lib.tal_bounds_masks(
            _ptr(self.l), _ptr(self.u), self.q, int(first_constraint), N, _ptr(op_safe),
            _ptr(pess), _ptr(opt), _ptr(pm), _ptr(pe), _ptr(max_l), _stream()),
            "tal_bounds_masks")
        return pess, opt, pm, pe, max_l

    def mask_to_indices(self, mask: torch.Tensor):
        """(idx[N] int64 (first `count` valid, ascending), count[1] int64 device)."""
        idx = torch.empty((self.N,), dtype=torch.int64, device=self.device)
        count = torch.empty((1,), dtype=torch.int64, device=self.device)
        _abi.check(self.lib.tal_mask_to_indices(_ptr(mask), mask.numel(), _ptr(idx), _ptr(count),
                                                _stream()), "tal_mask_to_indices")
        return idx, count

    # ----------------------------------------------------------------- scoring
    def score_itl_undirected(self) -> torch.Tensor:
        F = torch.empty((self.N,), dtype=torch.float64, device=self.device)
        _abi.check(self.lib.tal_score_itl_undirected(_ptr(self.var), _ptr(self.rho), self.q,
                                                     self.N, _ptr(F), _stream()),
                   "tal_score_itl_undirected")
        return F

    def _n_split(self, m: int) -> int:
        col_tiles = max(1, self.ld * (8 if self.dtype == torch.float64 else 4) // 4096)
        return int(max(1, min(64, m, -(-1184 // col_tiles))))

    def score_rows(self, rule: int, roi_idx: torch.Tensor, m: int, p0=None, p1: float = 0.0,
                   sharded_cols: bool = False, roi_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """F over this shard's candidates for the row-reduce rules (VTL, CTL,
        MM-ITL, TruVar).  One GPU: F is [N]; row-sharded with `sharded_cols`:
        F is [n_loc] for candidates row0..row0+n_loc (symmetric column gather)."""
        lib = self.lib
        N = self.N
        if self.lower and self._pending == 1 and 4 * m >= N and os.environ.get("TAL_FUSED_VTL", "1") != "0":
            # a VTL-type step: ONE pending update and an ROI that covers most of the domain -- the update pass and
            # both halves of the score in one transfer of the matrix (tal_update_colsum_lower_dense)
            return self._score_rows_lower(rule, roi_idx, m, p0, p1, roi_mask, fuse_update=True)
        self.flush()
        if self.lower:
            return self._score_rows_lower(rule, roi_idx, m, p0, p1, roi_mask)
        sharded_cols = sharded_cols or self.sharded
        n_out = self.n_loc if sharded_cols else N
        off = self.row0 if sharded_cols else 0
        F = torch.empty((n_out,), dtype=torch.float64, device=self.device)
        c = self._scratch("score_c", (n_out,))
        tr = self._scratch("score_tr", (1,))
        need_w = rule != _abi.RULE_VTL
        need_dinv = rule in (_abi.RULE_MMITL, _abi.RULE_TRUVAR)
        w = self._scratch("score_w", (max(m, 1),)) if need_w else None
        dinv = self._scratch("score_dinv", (N,)) if need_dinv else None
        n_split = self._n_split(m)
        part = self._scratch("score_part", (n_split, N)) if not sharded_cols else None
        if m == 0:
            F.zero_()
            if rule == _abi.RULE_VTL:
                F.neg_()
            return F
        for i in range(self.q):
            if need_w:
                _abi.check(lib.tal_roi_weights(_ptr(self.var[i]), _ptr(roi_idx), m,
                                               0 if rule == _abi.RULE_TRUVAR else 1, _ptr(w),
                                               _stream()), "tal_roi_weights")
            if need_dinv:
                _abi.check(lib.tal_pred_var_inv(_ptr(self.var[i]), _ptr(self.rho[i]), N,
                                                _ptr(dinv), _stream()), "tal_pred_var_inv")
            pp0 = float(p0[i]) if p0 is not None else 0.0
            if sharded_cols:
                _abi.check(lib.tal_colsum_cols(
                    rule, _ptr(self._cov[i]), self.ld, self.n_loc, _ptr(roi_idx), m, _ptr(w),
                    _ptr(dinv[off:]) if dinv is not None else None, pp0, float(p1), _ptr(c),
                    self.dt, _stream()), "tal_colsum_cols")
            else:
                _abi.check(lib.tal_colsum_rows(
                    rule, _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, N, _ptr(roi_idx), m,
                    _ptr(w), _ptr(dinv), pp0, float(p1), _ptr(part), n_split, _ptr(c), self.dt,
                    _stream()), "tal_colsum_rows")
            if rule == _abi.RULE_VTL:
                _abi.check(lib.tal_roi_trace(_ptr(self.var[i]), _ptr(roi_idx), m, _ptr(tr),
                                             _stream()), "tal_roi_trace")
            _abi.check(lib.tal_score_rows_finish(
                rule, _ptr(c), _ptr(self.var[i, off:]), _ptr(self.rho[i, off:]), _ptr(tr), n_out,
                1 if i > 0 else 0, _ptr(F), _stream()), "tal_score_rows_finish")
        return F

    def _score_rows_lower(self, rule: int, roi_idx: torch.Tensor, m: int, p0=None, p1: float = 0.0,
                          roi_mask: Optional[torch.Tensor] = None, fuse_update: bool = False) -> torch.Tensor:
        """The row-reduce rules in lower-block storage: every pair (a, x) of a target point and a
        candidate is kept either in row a (x below the end of a's diagonal block) or in row x;
        `tal_colsum_lower` sums what this shard keeps over all N candidates, a sum over the shards
        completes the vector, and every rank finishes its own candidates."""
        lib, N = self.lib, self.N
        n_out, off = self.n_cand, self.cand0
        F = torch.empty((n_out,), dtype=torch.float64, device=self.device)
        if m == 0:
            F.zero_()
            if rule == _abi.RULE_VTL:
                F.neg_()
            return F
        c = self._scratch("score_c_full", (N,))
        tr = self._scratch("score_tr", (1,))
        need_w = rule != _abi.RULE_VTL
        need_dinv = rule in (_abi.RULE_MMITL, _abi.RULE_TRUVAR)
        w = self._scratch("score_w", (max(m, 1),)) if need_w else None
        dinv = self._scratch("score_dinv", (N,)) if need_dinv else None
        # an ROI that covers a large part of the domain (m ~ N for the potential maximisers of the Safe-BO
        # configurations): stream + mask the rows instead of gathering them entry by entry; row splits of ~256
        # rows (the triangular kept part makes the CTAs unequal: many small ones keep the tail short)
        dense = 4 * m >= N
        n_split = int(max(1, min(1024, -(-self.n_loc // 256)))) if dense else self._n_split(m)
        part = self._scratch("score_part", (n_split, N))
        if dense:
            if roi_mask is None or roi_mask.numel() != N:
                roi_mask = torch.zeros((N,), dtype=torch.uint8, device=self.device)
                roi_mask[roi_idx[:m]] = 1
            roi_mask = roi_mask.view(torch.uint8) if roi_mask.dtype == torch.bool else roi_mask
            w_pt = self._scratch("score_w_pt", (N,)) if need_w else None
        for i in range(self.q):
            if need_w:
                _abi.check(lib.tal_roi_weights(_ptr(self.var[i]), _ptr(roi_idx), m,
                                               0 if rule == _abi.RULE_TRUVAR else 1, _ptr(w),
                                               _stream()), "tal_roi_weights")
                if dense:
                    w_pt.zero_()
                    w_pt[roi_idx[:m]] = w[:m]
            if need_dinv:
                _abi.check(lib.tal_pred_var_inv(_ptr(self.var[i]), _ptr(self.rho[i]), N,
                                                _ptr(dinv), _stream()), "tal_pred_var_inv")
            pp0 = float(p0[i]) if p0 is not None else 0.0
            if fuse_update:
                rowpart = self._scratch("score_rowpart", (int(lib.tal_update_colsum_scratch_bytes(
                    self.ld, self.n_loc, self.dt)) // 8,))
                _abi.check(lib.tal_update_colsum_lower_dense(
                    rule, _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, N, _ptr(self._Kp[0, i]),
                    _ptr(self._Wp[0, i]), _ptr(roi_mask), _ptr(w_pt), _ptr(dinv), pp0, float(p1), _ptr(part),
                    n_split, _ptr(rowpart), _ptr(c), self.dt_upd, _stream()), "tal_update_colsum_lower_dense")
            elif dense:
                _abi.check(lib.tal_colsum_lower_dense(
                    rule, _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, N, _ptr(roi_idx), m, _ptr(roi_mask),
                    _ptr(w), _ptr(w_pt), _ptr(dinv), pp0, float(p1), _ptr(part), n_split, _ptr(c), self.dt,
                    _stream()), "tal_colsum_lower_dense")
            else:
                _abi.check(lib.tal_colsum_lower(
                    rule, _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, N, _ptr(roi_idx), m, _ptr(w),
                    _ptr(dinv), pp0, float(p1), _ptr(part), n_split, _ptr(c), self.dt, _stream()),
                    "tal_colsum_lower")
            self._sum(c)
            if rule == _abi.RULE_VTL:
                _abi.check(lib.tal_roi_trace(_ptr(self.var[i]), _ptr(roi_idx), m, _ptr(tr),
                                             _stream()), "tal_roi_trace")
            _abi.check(lib.tal_score_rows_finish(
                rule, _ptr(c[off:]), _ptr(self.var[i, off:]), _ptr(self.rho[i, off:]), _ptr(tr), n_out,
                1 if i > 0 else 0, _ptr(F), _stream()), "tal_score_rows_finish")
        if fuse_update:  # the pending update went into the matrix with the pass above
            self._pending = 0
            self._upper_stale = True
        return F

    def _cond_dims(self, m: int):
        return round_up(m, 128), round_up(max(self.n_cand, 1), 64)

    def cond_var_sumsq(self, i: int, roi_idx: torch.Tensor, m: int, jitter: float = 1e-5,
                       V: Optional[torch.Tensor] = None,
                       cs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cs[x] = || L^-1 cov_i[A, x] ||^2 with L = chol(cov_i[A, A] + jitter^2 I) for
        this shard's candidates x: the part of compute_posterior_covariance's
        diagonal that directed ITL uses (itl.py:41-48).  V / cs: optional
        persistent buffers ([m_pad, ncols] / [ncols]) that receive L^-1 cov_i[A, :]
        and cs.  Row-sharded: cov[A, x] is read as the symmetric cov[x, A], Sigma_AA
        is assembled from the owners' rows by a sum, its Cholesky is replicated."""
        lib = self.lib
        self.flush()
        m_pad, ncols = self._cond_dims(m)
        ldb = ncols
        S = self._scratch("cv_S", (m_pad, m_pad))
        B = self._scratch("cv_B", (m_pad, ldb)) if V is None else V
        work = self._scratch("cv_work", (int(lib.tal_potrf_work_bytes(m_pad)) // 8,))
        if self.lower and self.comm is not None:
            self._gather_roi_lower_sharded(i, roi_idx, m, m_pad, ncols, jitter, S, B)
        elif self.comm is None:
            gather = lib.tal_gather_roi_lower if self.lower else lib.tal_gather_roi
            _abi.check(gather(_ptr(self._cov[i]), self.ld, self.N, _ptr(roi_idx), m,
                                          m_pad, float(jitter), _ptr(S), m_pad, _ptr(B), ldb,
                                          self.dt, _stream()), "tal_gather_roi")
        else:
            _abi.check(lib.tal_gather_roi_sharded(
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, _ptr(roi_idx), m, m_pad,
                float(jitter), _ptr(S), m_pad, _ptr(B), ldb, ncols, self.dt, _stream()),
                "tal_gather_roi_sharded")
            self._sum(S)
        _abi.check(lib.tal_potrf_lower(_ptr(S), m_pad, m_pad, _ptr(work), _stream()),
                   "tal_potrf_lower")
        _abi.check(lib.tal_trsm_lower(_ptr(S), m_pad, m_pad, _ptr(work), _ptr(B), ldb, ncols,
                                      _stream()), "tal_trsm_lower")
        n_split = int(max(1, min(64, m_pad // 64, -(-1184 // max(1, ncols // 512)))))
        part = self._scratch("cv_part", (n_split, ncols))
        if cs is None:
            cs = self._scratch("cv_cs", (ncols,))
        _abi.check(lib.tal_colsumsq_dense(_ptr(B), ldb, m_pad, ncols, _ptr(part), n_split,
                                          _ptr(cs), _stream()), "tal_colsumsq_dense")
        return cs[: self.n_cand]

    def _gather_roi_lower_sharded(self, i, roi_idx, m, m_pad, ncols, jitter, S, B) -> None:
        """Sigma_AA (+ jitter^2 I, identity pad) into S and Sigma[A, own candidates] into B for a
        row-sharded model in lower-block storage: the rows of the target set are assembled
        64 at a time by the symmetric row gather + a sum over the shards (one-off per ROI)."""
        S.zero_()
        B.zero_()
        rows = self._scratch("cv_rows", (64, self.ld))
        idx = roi_idx[:m].contiguous()
        for j0 in range(0, m, 64):
            j1 = min(m, j0 + 64)
            _abi.check(self.lib.tal_gather_rows_lower(
                _ptr(self._cov[i]), self.ld, self.row0, self.n_loc, self.N, _ptr(idx[j0:j1]), j1 - j0,
                _ptr(rows), self.ld, self.dt, _stream()), "tal_gather_rows_lower")
            self._sum(rows)
            B[j0:j1, : self.n_cand] = rows[: j1 - j0, self.cand0: self.cand0 + self.n_cand]
            S[j0:j1, :m] = rows[: j1 - j0].index_select(1, idx)
        d = torch.arange(m_pad, device=self.device)
        S[d[:m], d[:m]] += float(jitter) ** 2
        S[d[m:], d[m:]] = 1.0

    def score_itl_directed(self, roi_idx: torch.Tensor, m: int, jitter: float = 1e-5,
                           incremental: bool = False, roi_key=None) -> torch.Tensor:
        """sum_i 0.5 max(log((var_i + rho_i^2)/(var_i - cs_i + rho_i^2)), 0) (itl.py:50-58)
        over this shard's candidates ([N] on one GPU, [n_cand] when row-sharded).
        incremental: keep V_i = L^-1 cov_i[A, :] resident and let `observe`
        downdate it (csrc/condinc.cu) as long as `roi_key` stays the same; the
        full Cholesky + TRSM then only runs when the ROI changes."""
        n_out, off = self.n_cand, self.cand0
        F = torch.empty((n_out,), dtype=torch.float64, device=self.device)
        st = self._cond
        reuse = incremental and st is not None and st["key"] is roi_key and st["valid"]
        if not reuse:
            self._sel_discard()
        if incremental and not reuse:
            m_pad, ncols = self._cond_dims(m)
            append = self.cond_form == "append"
            cap = m_pad + (2 * self.cond_capacity if append else 0)
            if st is None or st["V"].shape != (self.q, cap, ncols):
                self._cond = st = None  # drop the old buffers before allocating
                f64 = dict(dtype=torch.float64, device=self.device)
                st = dict(V=torch.zeros((self.q, cap, ncols), **f64), cs=torch.empty((self.q, ncols), **f64),
                          pcol=torch.zeros((cap,), **f64))
                if append:
                    blocks = max(1, ncols // 512)
                    st["n_split"] = int(max(1, min(64, m_pad // 64, -(-1184 // blocks))))
                    st["app"] = torch.empty((int(self.lib.tal_cond_append_scratch_bytes(ncols, st["n_split"])) // 8,),
                                            **f64)
                    st["fz"] = torch.zeros((int(self.lib.tal_fused_scratch_bytes(self.q, ncols, st["n_split"])),),
                                           dtype=torch.uint8, device=self.device)
                else:
                    st["coef"] = torch.empty((int(self.lib.tal_cond_coef_bytes(m_pad)) // 8,), **f64)
                    st["chunk"] = torch.empty((int(self.lib.tal_cond_chunk_scratch_bytes(ncols)) // 8,), **f64)
            st.update(key=roi_key, m_pad=m_pad, ncols=ncols, cap=cap, rows=m_pad, append=append, valid=False)
            self._cond = st
            self._gen += 1
            if self._peer is not None and self._peer.m_pad < cap:
                self._peer_attach(cap)
        if not incremental and st is not None:
            st["valid"] = False
        for i in range(self.q):
            if reuse:
                cs = st["cs"][i, :n_out]
            elif incremental:
                cs = self.cond_var_sumsq(i, roi_idx, m, jitter, V=st["V"][i, : st["m_pad"]], cs=st["cs"][i])
            else:
                cs = self.cond_var_sumsq(i, roi_idx, m, jitter)
            _abi.check(self.lib.tal_score_itl_directed(
                _ptr(self.var[i, off:]), _ptr(cs), _ptr(self.rho[i, off:]), n_out,
                1 if i > 0 else 0, _ptr(F), _stream()), "tal_score_itl_directed")
        if incremental:
            st["valid"] = True
        return F

    def _cond_downdate(self, idx_dev, idx_host, allreduce=None) -> None:
        """After exchange_row + step_vectors of an observation, carry the target-set
        conditioning along: append form (csrc/condapp.cu) or V_i <- T^-1 (V_i - V_i[:, x] w^T)
        (csrc/condinc.cu)."""
        st = self._cond
        rows = st["rows"]
        if st["append"] and rows + 2 > st["cap"]:
            st["valid"] = False  # capacity reached: the next scoring pass re-conditions from scratch
            return
        for i in range(self.q):
            V = st["V"][i]
            if self._peer is not None:  # the owner stored V[i][:, x] with the row (exchange_row)
                pcol = self._peer.pcol[i]
            else:
                pcol = st["pcol"]
                _abi.check(self.lib.tal_cond_extract_col(
                    _ptr(V), st["ncols"], rows, self.cand0, self.n_cand, _ptr(idx_dev),
                    int(idx_host), _ptr(pcol), _stream()), "tal_cond_extract_col")
                self._sum(pcol[:rows])
            if st["append"]:
                _abi.check(self.lib.tal_cond_append_update(
                    _ptr(V), st["ncols"], st["cap"], st["m_pad"], rows, st["ncols"], self.n_cand, self.cand0,
                    _ptr(pcol), _ptr(self.kbuf[i]), _ptr(self.wbuf[i]), _ptr(self.rho[i]), _ptr(idx_dev),
                    int(idx_host), _ptr(st["cs"][i]), _ptr(st["app"]), st["n_split"], self.dt, _stream()),
                    "tal_cond_append_update")
            else:
                _abi.check(self.lib.tal_cond_update(
                    _ptr(V), st["ncols"], st["m_pad"], st["ncols"], self.n_cand, self.cand0,
                    _ptr(pcol), _ptr(self.kbuf[i]), _ptr(self.wbuf[i]), _ptr(self.rho[i]),
                    _ptr(idx_dev), int(idx_host), _ptr(st["coef"]), _ptr(st["cs"][i]),
                    int(self.cond_chunks), _ptr(st["chunk"]), self.dt, _stream()), "tal_cond_update")
        if st["append"]:
            st["rows"] = rows + 2

    # ------------------------------------------- dense blocks
    def _need_unsharded(self, what: str) -> None:
        if self.comm is not None:
            raise NotImplementedError(f"{what} factors a dense block of the covariance and needs all of its rows "
                                      "on one GPU (row-sharded models: not built)")

    def block_cholesky(self, i: int, idx: Optional[torch.Tensor], m: int, add_std: Optional[torch.Tensor] = None,
                       add: float = 0.0, name: str = "blk"):
        """L = chol(cov_i[idx, idx] + diag(add + add_std^2)) on the DMMA path; returns (S [m_pad, m_pad]
        with L in its lower triangle, m_pad).  idx None: the leading m x m block."""
        self._need_unsharded("block_cholesky")
        self.flush()
        m_pad = round_up(max(m, 1), 128)
        S = self._scratch(name + "_S", (m_pad, m_pad))
        work = self._scratch("cv_work", (int(self.lib.tal_potrf_work_bytes(m_pad)) // 8,))
        _abi.check(self.lib.tal_gather_sym_block(
            _ptr(self._cov[i]), self.ld, self.N, _ptr(idx), m, m_pad, _ptr(add_std), float(add), _ptr(S), m_pad,
            int(self.lower), self.dt, _stream()), "tal_gather_sym_block")
        _abi.check(self.lib.tal_potrf_lower(_ptr(S), m_pad, m_pad, _ptr(work), _stream()), "tal_potrf_lower")
        return S, m_pad

    def chol_logdet(self, S: torch.Tensor, m_pad: int, m: int) -> torch.Tensor:
        """slogdet of the factored block: 2 sum log diag(L) (device scalar)."""
        out = torch.empty((1,), dtype=torch.float64, device=self.device)
        _abi.check(self.lib.tal_chol_logdet(_ptr(S), m_pad, m, _ptr(out), _stream()), "tal_chol_logdet")
        return out[0]

    def block_logdet(self, i: int, idx: Optional[torch.Tensor], m: int, add_std=None, add: float = 0.0) -> torch.Tensor:
        if m == 0:
            return torch.zeros((), dtype=torch.float64, device=self.device)
        S, m_pad = self.block_cholesky(i, idx, m, add_std, add)
        return self.chol_logdet(S, m_pad, m)

    def sample(self, i: int, z: torch.Tensor, rel_jitter: float = 1e-8) -> torch.Tensor:
        """mean_i + L z for k draws z [k, N] with L L^T = cov_i + jitter I: one Cholesky and one pass over L
        for every 8 draws.  jitter = rel_jitter * max(var), raised 100x (up to three times) while the factor
        is not finite (the reference's SVD accepts a numerically indefinite matrix, Cholesky does not)."""
        N = self.N
        z = z.to(device=self.device, dtype=torch.float64).contiguous()
        k = z.shape[0]
        scale = float(self.var[i].max())
        for attempt in range(4):
            S, m_pad = self.block_cholesky(i, None, N, add=rel_jitter * scale * (100.0 ** attempt), name="smp")
            if bool(torch.isfinite(torch.diagonal(S)[:N]).all()):
                break
        out = torch.empty((k, N), dtype=torch.float64, device=self.device)
        _abi.check(self.lib.tal_tri_sample(_ptr(S), m_pad, N, _ptr(z), N, k, _ptr(self.mean[i]), _ptr(out), N,
                                           _stream()), "tal_tri_sample")
        return out

    def information_gain(self, i: int, roi_idx: torch.Tensor, m: int, obs_idx: torch.Tensor, nb: int,
                         obs_noise_std: Optional[torch.Tensor], jitter: float = 1e-5) -> torch.Tensor:
        """0.5 max(logdet(Sigma_BB + D) - logdet(Sigma_BB - Sigma_BA (Sigma_AA + jitter^2 I)^-1 Sigma_AB + D), 0)
        (GaussianDistribution.information_gain, lib/gp/gaussian_distribution.py:151-180): gather -> Cholesky ->
        forward TRSM (the conditioning kernels of directed ITL) -> SYRK of the columns B -> two Choleskys."""
        self._need_unsharded("information_gain")
        lib = self.lib
        obs_idx = obs_idx.to(device=self.device, dtype=torch.int64).contiguous()
        if obs_noise_std is not None:
            obs_noise_std = obs_noise_std.to(device=self.device, dtype=torch.float64).contiguous()
        add = 0.0  # noise_covariance_matrix(k, None) has default = 0 (lib/utils.py:44-50)
        prior_ld = self.block_logdet(i, obs_idx, nb, obs_noise_std, add)
        nb_pad = round_up(nb, 128)
        P = self._scratch("ig_P", (nb_pad, nb_pad))
        _abi.check(lib.tal_gather_sym_block(_ptr(self._cov[i]), self.ld, self.N, _ptr(obs_idx), nb, nb_pad,
                                            _ptr(obs_noise_std), add, _ptr(P), nb_pad, int(self.lower), self.dt,
                                            _stream()), "tal_gather_sym_block")
        if m > 0:
            m_pad, ncols = self._cond_dims(m)
            V = self._scratch("cv_B", (m_pad, ncols))
            self.cond_var_sumsq(i, roi_idx, m, jitter, V=V)
            Vt = self._scratch("ig_Vt", (nb_pad, m_pad))
            _abi.check(lib.tal_gather_cols_t(_ptr(V), ncols, m_pad, _ptr(obs_idx), nb, nb_pad, _ptr(Vt), m_pad,
                                             _stream()), "tal_gather_cols_t")
            _abi.check(lib.tal_syrk_lower_sub(_ptr(P), nb_pad, nb_pad, _ptr(Vt), m_pad, m_pad, _stream()),
                       "tal_syrk_lower_sub")
        work = self._scratch("cv_work", (int(lib.tal_potrf_work_bytes(nb_pad)) // 8,))
        _abi.check(lib.tal_potrf_lower(_ptr(P), nb_pad, nb_pad, _ptr(work), _stream()), "tal_potrf_lower")
        post_ld = torch.nan_to_num(self.chol_logdet(P, nb_pad, nb), nan=-np.inf)  # :175-177
        return 0.5 * torch.clamp(prior_ld - post_ld, min=0.0)

    def drop_conditioning(self) -> None:
        self._sel_discard()
        self._gen += 1
        self._cond = None
        self._ctx = None  # (its cached pointers into the conditioning buffers are gone)

    # ------------------------------------------------------------------ argmax
    def masked_argmax(self, F: torch.Tensor, safe: Optional[torch.Tensor], first_constraint: int,
                      sample: Optional[torch.Tensor], idx_offset: int = 0,
                      l_offset: int = 0) -> torch.Tensor:
        """Enqueue the masked arg-max; returns the device result record
        (int64[4] view of tal_argmax_result)."""
        n = F.numel()
        _abi.check(self.lib.tal_masked_argmax(
            _ptr(F), _ptr(safe), _ptr(self.l[:, l_offset:]) if safe is None else None, self.q,
            int(first_constraint), self.N, _ptr(sample), n, int(idx_offset),
            _ptr(self._argmax_scratch), _ptr(self._result), _stream()), "tal_masked_argmax")
        return self._result

    def argmax_candidates(self, F_loc: torch.Tensor, safe: Optional[torch.Tensor],
                          first_constraint: int, sample: Optional[torch.Tensor]) -> torch.Tensor:
        """Masked arg-max over ALL candidates given this shard's scores: local
        reduce, all-gather of the 32-byte records, identical combine on every
        rank.  `safe` / `sample` are full-length [N] masks (or None).  Returns
        the device record (int64[4]; [0] is the global index)."""
        if self.comm is None:
            return self.masked_argmax(F_loc, safe, first_constraint, sample)
        sl = slice(self.cand0, self.cand0 + self.n_cand)
        rec = self.masked_argmax(F_loc, safe[sl] if safe is not None else None, first_constraint,
                                 sample[sl] if sample is not None else None, idx_offset=self.cand0,
                                 l_offset=self.cand0)
        self.combine_records(rec)
        return self._combined

    def combine_records(self, rec: torch.Tensor) -> None:
        """_combined <- lexicographic reduce of every shard's arg-max record."""
        comm = self.comm
        if self._peer is not None:
            self._sel_discard()  # (a record published ahead of time by the fused step is consumed first)
            _abi.check(self.lib.tal_peer_publish_record(_ptr(rec), comm.table_ptr, comm.world, comm.rank,
                                                        _stream()), "tal_peer_publish_record")
            _abi.check(self.lib.tal_peer_combine_records(comm.box_ptr, comm.world, _ptr(self._combined),
                                                         _stream()), "tal_peer_combine_records")
            return
        comm.gather_records(rec, self._recs)
        _abi.check(self.lib.tal_argmax_combine(_ptr(self._recs), comm.world,
                                               _ptr(self._combined), _stream()), "tal_argmax_combine")

    def read_result(self, result: Optional[torch.Tensor] = None) -> _abi.ArgmaxResult:
        """Device -> host read of an arg-max record (synchronises the stream)."""
        r = self._result if result is None else result
        self._result_host.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = _abi.ArgmaxResult.from_buffer_copy(self._result_host.numpy().tobytes())
        if res.status == _abi.ARGMAX_PEER_ERROR or (self._peer is not None and res.flags & 8):
            raise _abi.TalError(f"NVLink peer exchange timed out (TAL_PEER_ERR {self.comm.error()}): "
                                "a rank is late or desynchronised; the replicated state is no longer trustworthy")
        return res
