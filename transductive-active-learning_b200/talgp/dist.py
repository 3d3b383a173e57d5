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
def shard_bounds(N: int, world: int, align: int = 64, storage: str = "full") -> List[int]:
    """Row boundaries b[0..world]: rank g holds global rows [b[g], b[g+1]).
    full storage: equal blocks (multiples of `align` rows except the last);
    lower-block storage: the kept part of row r is ~r columns long, so the
    boundaries sit at N sqrt(g / world) to balance the update traffic."""
    if storage == "lower":
        # exact cost model: row r keeps cend(r) = min(ld, (r // B + 1) * B) columns, B = tal_sym_block = 32
        B, ld = 32, -(-N // 32) * 32
        cend = np.minimum((np.arange(N, dtype=np.int64) // B + 1) * B, ld)
        cum = np.concatenate([[0], np.cumsum(cend)])
        b = [0]
        for g in range(1, world):
            r = int(np.searchsorted(cum, cum[-1] * g / world))
            b.append(min(N, max(b[-1], int(round(r / align)) * align)))
        return b + [N]
    per = -(-N // world)
    per = -(-per // align) * align
    return [min(N, g * per) for g in range(world)] + [N]


def shard_rows(N: int, world: int, rank: int, align: int = 64, storage: str = "full") -> Tuple[int, int]:
    """Contiguous row block of `rank`: (row0, n_loc)."""
    b = shard_bounds(N, world, align, storage)
    return b[rank], b[rank + 1] - b[rank]


def owner_of(N: int, world: int, idx: int, align: int = 64, storage: str = "full") -> int:
    b = shard_bounds(N, world, align, storage)
    for g in range(world):
        if b[g] <= int(idx) < b[g + 1]:
            return g
    raise IndexError(idx)


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
        sh["barrier"].wait(120)
        tot = sh["slots"][0].clone()
        for r in range(1, self.world):
            tot += sh["slots"][r]
        sh["barrier"].wait(120)  # every thread has issued its reads
        buf.copy_(tot)
        sh["barrier"].wait(120)
        return buf

    def gather_records(self, rec: torch.Tensor, out: torch.Tensor) -> None:
        sh = self.shared
        sh["slots"][self.rank] = rec
        sh["barrier"].wait(120)
        for r in range(self.world):
            out[r].copy_(sh["slots"][r])
        sh["barrier"].wait(120)

    def all_gather_(self, part: torch.Tensor, out: torch.Tensor) -> None:
        sh = self.shared
        sh["slots"][self.rank] = part
        sh["barrier"].wait(120)
        n = part.numel()
        for r in range(self.world):
            out[r * n:(r + 1) * n].copy_(sh["slots"][r])
        sh["barrier"].wait(120)


# ------------------------------------------------------ NVLink peer-memory exchange
class _DevMem:
    """A raw device allocation exposed through __cuda_array_interface__ so that
    torch can alias it without owning it."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    ts = {torch.float64: "<f8", torch.float32: "<f4", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
    return torch.as_tensor(_DevMem(ptr, shape, ts), device=device)


class PeerViews:
    def __init__(self, kbuf, pcol, kbuf_off, pcol_off, m_pad):
        self.kbuf, self.pcol, self.kbuf_off, self.pcol_off, self.m_pad = kbuf, pcol, kbuf_off, pcol_off, m_pad


class _PeerMixin:
    """The per-step exchanges over NVLink peer memory (csrc/peer.cu): every rank
    owns a mailbox, the peers map it (CUDA IPC between processes, plain pointers
    between threads of one process), and the arg-max all-gather / observed-row
    broadcast become P2P stores + release/acquire sequence numbers.  Set-up
    collectives (handle exchange, Sigma_AA assembly) still use the base class."""
    peer = True

    def _peer_init(self):
        from . import _abi
        self.lib = _abi.load()
        self._box = None        # own mailbox (device pointer, int)
        self._peers = []        # mailbox pointers as seen from this rank
        self._imported = []
        self._table = None      # device array of the G pointers
        self._layout = None

    def _exchange_ptrs(self, own_ptr: int):
        raise NotImplementedError

    def _host_barrier(self):
        raise NotImplementedError

    def attach(self, q: int, ld: int, dtype: torch.dtype, m_pad: int, device) -> PeerViews:
        """(Re)allocate the mailboxes for a [q, ld] row buffer and a [q, m_pad]
        conditioning column.  Collective: every rank calls it at the same point."""
        from . import _abi
        lib = self.lib
        if self.world > _abi.TAL_PEER_MAX:
            raise ValueError("too many ranks for the peer mailbox")
        item = 8 if dtype == torch.float64 else 4
        up = lambda x: (x + 255) // 256 * 256  # noqa: E731
        hb = int(lib.tal_peer_header_bytes())
        kbuf_off = hb
        pcol_off = kbuf_off + up(q * ld * item)
        total = pcol_off + up(max(1, q * m_pad) * 8)
        torch.cuda.synchronize(device)
        self._host_barrier()
        self.release()
        box = C.c_void_p()
        _abi.check(lib.tal_peer_alloc(total, C.byref(box)), "tal_peer_alloc")
        self._box = int(box.value)
        self._peers = self._exchange_ptrs(self._box)
        self._table = torch.tensor(self._peers, dtype=torch.int64, device=device)
        self._layout = (q, ld, dtype, m_pad)
        torch.cuda.synchronize(device)
        self._host_barrier()
        kbuf = _view(self._box + kbuf_off, (q, ld), dtype, device)
        pcol = _view(self._box + pcol_off, (q, max(1, m_pad)), torch.float64, device)
        return PeerViews(kbuf, pcol, kbuf_off, pcol_off, m_pad)

    def release(self):
        from . import _abi
        if self._box is None:
            return
        for p in self._imported:
            _abi.check(self.lib.tal_peer_unimport(C.c_void_p(p)), "tal_peer_unimport")
        self._imported = []
        self._host_barrier()  # nobody maps this rank's mailbox any more
        _abi.check(self.lib.tal_peer_free(C.c_void_p(self._box)), "tal_peer_free")
        self._box, self._peers, self._table = None, [], None

    @property
    def table_ptr(self):
        return C.c_void_p(self._table.data_ptr())

    @property
    def box_ptr(self):
        return C.c_void_p(self._box)

    def error(self) -> int:
        from . import _abi
        err = C.c_int(0)
        _abi.check(self.lib.tal_peer_error(self.box_ptr, C.byref(err),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tal_peer_error")
        return int(err.value)


class PeerComm(_PeerMixin, NcclComm):
    """One process per GPU (torchrun): NCCL for set-up, CUDA-IPC peer memory for
    the per-step exchanges."""

    def __init__(self, group=None):
        NcclComm.__init__(self, group)
        self._peer_init()

    def _host_barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def _exchange_ptrs(self, own_ptr: int):
        from . import _abi
        h = C.create_string_buffer(_abi.TAL_PEER_HANDLE_BYTES)
        _abi.check(self.lib.tal_peer_export(C.c_void_p(own_ptr), h), "tal_peer_export")
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(h.raw), group=self.group)
        else:
            handles = [bytes(h.raw)]
        ptrs = []
        for r, hb in enumerate(handles):
            if r == self.rank:
                ptrs.append(own_ptr)
                continue
            p = C.c_void_p()
            _abi.check(self.lib.tal_peer_import(hb, C.byref(p)), "tal_peer_import")
            self._imported.append(int(p.value))
            ptrs.append(int(p.value))
        return ptrs


class LocalPeerComm(_PeerMixin, ThreadComm):
    """The same peer-memory protocol between `world` threads of ONE process that
    drive one shard each on the same GPU, every thread on its OWN stream (the
    kernels of different shards must run concurrently: they wait on each other).
    Pointers are shared directly; used to test csrc/peer.cu on a single GPU."""

    single_context = True  # every shard lives in ONE CUDA context: see DeviceGP._ctx_get (no side stream)

    def __init__(self, shared: dict, rank: int, world: int):
        ThreadComm.__init__(self, shared, rank, world)
        self._peer_init()

    def _host_barrier(self):
        self.shared["barrier"].wait(120)

    def _exchange_ptrs(self, own_ptr: int):
        sh = self.shared
        sh["slots"][self.rank] = own_ptr
        sh["barrier"].wait(120)
        ptrs = [int(sh["slots"][r]) for r in range(self.world)]
        sh["barrier"].wait(120)
        return ptrs

    # ThreadComm's collectives rely on one shared stream (issue order = execution
    # order); with one stream per thread every hand-over needs a stream sync.
    def sum_(self, buf):
        sh, st = self.shared, torch.cuda.current_stream()
        st.synchronize()
        sh["slots"][self.rank] = buf
        sh["barrier"].wait(120)
        tot = sh["slots"][0].clone()
        for r in range(1, self.world):
            tot += sh["slots"][r]
        st.synchronize()
        sh["barrier"].wait(120)  # every thread has finished reading
        buf.copy_(tot)
        st.synchronize()
        sh["barrier"].wait(120)
        return buf

    def all_gather_(self, part, out):
        sh, st = self.shared, torch.cuda.current_stream()
        st.synchronize()
        sh["slots"][self.rank] = part
        sh["barrier"].wait(120)
        n = part.numel()
        for r in range(self.world):
            out[r * n:(r + 1) * n].copy_(sh["slots"][r])
        st.synchronize()
        sh["barrier"].wait(120)

    def gather_records(self, rec, out):
        sh, st = self.shared, torch.cuda.current_stream()
        st.synchronize()
        sh["slots"][self.rank] = rec
        sh["barrier"].wait(120)
        for r in range(self.world):
            out[r].copy_(sh["slots"][r])
        st.synchronize()
        sh["barrier"].wait(120)


# ------------------------------------------------------------------ device side
def ShardedGP(N: int, q: int, dtype=torch.float64, device=None, comm=None, rank1_variant: int = 0,
              storage: str = "full"):
    """A DeviceGP holding this rank's row block (talgp.engine.DeviceGP with a communicator)."""
    from .engine import DeviceGP
    return DeviceGP(N, q, dtype=dtype, device=device, rank1_variant=rank1_variant,
                    comm=comm if comm is not None else default_comm(), storage=storage)


def default_comm(group=None):
    """PeerComm (NVLink peer-memory exchange) between the GPUs of one box when the
    process group runs on NCCL; NcclComm otherwise or with TAL_EXCHANGE=nccl."""
    import os
    if (dist.is_initialized() and dist.get_backend(group) == "nccl"
            and os.environ.get("TAL_EXCHANGE", "peer") != "nccl"):
        return PeerComm(group)
    return NcclComm(group)
