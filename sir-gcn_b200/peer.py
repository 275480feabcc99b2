"""Peer-memory transport of the row-partitioned graph (partition.py): all-gathers of the K / Q / dA row tables as
copy-engine pulls over NVLink instead of NCCL kernels (include/sirgcn.h, "Peer-memory transport").

Why: at 8 GPUs the three all-gathers of a layer move as many bytes over NVLink as the edge walks move through HBM
in the same time, so they must overlap the walks without slowing them.  An NCCL all-gather occupies SMs and
issues its loads/stores from them; a copy-engine pull takes no SM at all.

    slice = pool.acquire(rows, ld, dtype)      # this rank's [rows, ld] slice, IPC-mapped by every peer
    pool.ensure_writable(slice)
    ...producer kernels write slice.local (all of it, or rows lo..hi)...
    h = pool.gather(slice, full[, lo, hi])     # barrier ("rows written everywhere") + world-1 pulls on copy streams
    ...independent work...
    h.wait()                                   # the current stream waits for the pulls
    pool.release(slice)                        # may be handed out again

Hazards.  Every rank runs the same sequence of acquire / gather / wait / release (SPMD).  A slice may be overwritten
once every peer has finished pulling from it; a peer finishes its pulls before the `wait` it enqueues on its compute
stream, and a barrier starts only when the issuing compute stream has reached it — so ANY barrier issued after that
`wait` proves it, once the local compute stream has waited for that barrier (directly, or through the pulls of a later
gather, which run behind their barrier).  The pool numbers the barriers: a slice remembers the newest barrier issued
at its last `wait`, `joined` is the newest barrier the compute stream has waited for, and `ensure_writable` enqueues an
extra blocking barrier only when joined <= that number.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib

# a peer that has not arrived after this long is fatal: the barrier kernel records status 1 and traps, so nothing ever
# computes on tables that were not gathered (a rank stalled by a checkpoint save or an eval needs a longer limit:
# SIRGCN_BARRIER_TIMEOUT_S)
_BARRIER_TIMEOUT_NS = int(float(os.environ.get("SIRGCN_BARRIER_TIMEOUT_S", "120")) * 1e9)


class _RawDeviceBuffer:
    """exposes a raw device allocation through __cuda_array_interface__ so torch can view it"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerSlice:
    """this rank's slice of a gathered row table: `local` [rows, ld] lives in an IPC-exported allocation,
    `peer_ptr[r]` is rank r's slice mapped into this process (None for the own rank)"""

    def __init__(self, rows, ld, dtype, base, local, peer_ptr):
        self.rows, self.ld, self.dtype = rows, ld, dtype
        self.base, self.local, self.peer_ptr = base, local, peer_ptr
        self.row_bytes = ld * local.element_size()
        self.waited_at = -1         # index of the newest barrier issued when the last pull of this slice was waited for
        self.gathered = False       # peers have pulled (or may still be pulling) the current contents
        self.pending = []           # PullHandles of gathers that have not been waited for yet


class PullHandle:
    def __init__(self, pool, sl, events, barrier_index):
        self.pool, self.slice, self.events, self.barrier_index = pool, sl, events, barrier_index

    def wait(self):
        """the current stream waits for the pulls (the host does not block)"""
        if self.events is None:
            return
        cur = torch.cuda.current_stream(self.pool.device)
        for ev in self.events:
            cur.wait_event(ev)
        self.events = None
        pool, sl = self.pool, self.slice
        pool.joined = max(pool.joined, self.barrier_index)     # the pulls ran behind that barrier
        sl.waited_at = pool.barriers
        sl.pending = [h for h in sl.pending if h is not self]


class PeerFull:
    """a gathered table [world*rows, ld] in an IPC-exported allocation: every peer PUSHES its slice into its row range
    (`peer_ptr[r]` = rank r's table mapped into this process)"""

    def __init__(self, rows, ld, dtype, base, local, peer_ptr):
        self.rows, self.ld, self.dtype = rows, ld, dtype
        self.base, self.local, self.peer_ptr = base, local, peer_ptr
        self.row_bytes = ld * local.element_size()
        self.released_at = 0        # newest barrier issued when the last reader of this table was enqueued


class PushHandle:
    def __init__(self, pool, done, barrier_index):
        self.pool, self.done, self.barrier_index = pool, done, barrier_index

    def wait(self):
        """the current stream waits until every rank's rows have landed in the local table"""
        if self.done is not None:
            torch.cuda.current_stream(self.pool.device).wait_event(self.done)
            self.done = None
        self.pool.joined = max(self.pool.joined, self.barrier_index)


class PeerPool:
    """IPC-mapped row-table slices of one process group (one process per GPU of one node) + the pull all-gather"""

    def __init__(self, group=None, device=None, barrier="flags"):
        if not dist.is_initialized():
            raise RuntimeError("PeerPool needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.world > 32:
            raise ValueError("PeerPool covers the GPUs of one node (world <= 32)")
        self.lib = _lib.lib()
        self.free, self.all = {}, []
        self.barriers = 0           # barriers issued so far (the same count on every rank: SPMD)
        self.joined = 0             # newest barrier whose completion the compute stream has waited for
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, self.world - 1))]
        # every barrier runs on ONE stream of its own, so that barriers execute in issue order on every rank and a
        # gather's barrier does not make the compute stream wait for the slowest peer
        self.sync_stream = torch.cuda.Stream(device=self.device)
        self.push_stream = torch.cuda.Stream(device=self.device, priority=-1)   # SM-driven pushes: scheduled first
        self.push_ctas = int(os.environ.get("SIRGCN_PUSH_CTAS", "32"))
        self.free_full, self.all_full = {}, []
        self._last_done = (0, None)     # (index, done event) of the newest barrier
        self.barrier_kind = os.environ.get("SIRGCN_PEER_BARRIER", barrier)
        self._nccl_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        # flag pads: one uint32 per peer, in a peer allocation of their own
        self._pad_base, pad_peers = self._alloc_mapped(4096)
        pads = [self._pad_base if r == self.rank else pad_peers[r] for r in range(self.world)]
        self._pads = torch.tensor(pads, dtype=torch.int64, device=self.device)
        self._pad_peers = pad_peers
        self._epoch = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)      # every pad is mapped (and zero) before the first flag is written

    # ---- allocation ---------------------------------------------------------------------------------------------
    def _alloc_mapped(self, nbytes):
        """collective: (own device pointer, [peer r's allocation mapped here or None]).  A failure on ANY rank
        (allocation, IPC export or mapping) is agreed on by all ranks before anybody raises, so that the caller can
        fall back to another transport on every rank at once instead of leaving the others in a collective."""
        handle = (C.c_ubyte * 64)()
        base = C.c_void_p()
        err = None
        with torch.cuda.device(self.device):
            rc = self.lib.sirgcn_peer_alloc(C.c_size_t(nbytes), C.byref(base), handle)
        if rc != 0:
            err = f"sirgcn_peer_alloc failed (rc={rc}): {self.lib.sirgcn_last_error().decode()}"
        handles = [None] * self.world
        dist.all_gather_object(handles, None if err else bytes(handle), group=self.group)
        peers = []
        for r, h in enumerate(handles):
            if r == self.rank or h is None or err:
                peers.append(None)
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            with torch.cuda.device(self.device):
                rc = self.lib.sirgcn_peer_open(buf, C.byref(p))
            if rc != 0:
                err = f"sirgcn_peer_open failed (rc={rc}): {self.lib.sirgcn_last_error().decode()}"
                peers.append(None)
            else:
                peers.append(p.value)
        if any(h is None for h in handles) and not err:
            err = "a peer rank could not allocate or export its buffer"
        bad = torch.tensor([1 if err else 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=self.group)
        if int(bad.item()):
            for p in peers:
                if p:
                    self.lib.sirgcn_peer_close(C.c_void_p(p))
            dist.barrier(group=self.group)          # nobody frees memory a peer still has mapped
            if base.value:
                self.lib.sirgcn_peer_free(base)
            raise RuntimeError("peer-memory setup failed on at least one rank" + (f": {err}" if err else ""))
        return base.value, peers

    def acquire(self, rows, ld, dtype) -> PeerSlice:
        """collective the first time a (rows, ld, dtype) slice is needed; afterwards slices are recycled"""
        key = (rows, ld, dtype)
        lst = self.free.setdefault(key, [])
        if lst:
            return lst.pop()
        es = torch.empty((), dtype=dtype).element_size()
        nbytes = max(16, rows * ld * es)
        base, peers = self._alloc_mapped(nbytes)
        raw = torch.as_tensor(_RawDeviceBuffer(base, nbytes), device=self.device)
        local = raw.view(dtype)[:rows * ld].view(rows, ld)
        sl = PeerSlice(rows, ld, dtype, base, local, peers)
        self.all.append(sl)
        return sl

    def release(self, sl: PeerSlice):
        for h in list(sl.pending):          # never waited for (e.g. backward never ran): drain first
            h.wait()
        self.free.setdefault((sl.rows, sl.ld, sl.dtype), []).append(sl)

    # ---- ordering -----------------------------------------------------------------------------------------------
    def barrier(self, blocking=True, after=()):
        """device-side barrier of all ranks: it starts once the current stream has reached this point (and the events
        in `after` have fired) and runs on the pool's sync stream; with `blocking` the current stream also waits for
        it.  Returns (index, done event)."""
        self.barriers += 1
        index = self.barriers
        cur = torch.cuda.current_stream(self.device)
        if self.world == 1:
            self.joined = index
            return index, None
        here = torch.cuda.Event()
        here.record(cur)
        self.sync_stream.wait_event(here)
        for ev in after:
            self.sync_stream.wait_event(ev)
        if self.barrier_kind == "nccl":
            with torch.cuda.stream(self.sync_stream):
                dist.all_reduce(self._nccl_flag, group=self.group)
        else:
            self._epoch += 1
            with torch.cuda.device(self.device):
                rc = self.lib.sirgcn_peer_barrier(C.c_void_p(self._pads.data_ptr()), C.c_int32(self.world),
                                                  C.c_int32(self.rank), C.c_uint32(self._epoch & 0xffffffff),
                                                  C.c_uint64(_BARRIER_TIMEOUT_NS), C.c_void_p(self._status.data_ptr()),
                                                  C.c_void_p(self.sync_stream.cuda_stream))
            _lib.check(rc, "sirgcn_peer_barrier")
        done = torch.cuda.Event()
        done.record(self.sync_stream)
        self._last_done = (index, done)
        if blocking:
            cur.wait_event(done)
            self.joined = index
        return index, done

    def check(self):
        """host-side: raise if any barrier gave up waiting for a peer (synchronises the device)"""
        code = int(self._status.item())
        if code != 0:
            raise RuntimeError("sirgcn_peer_barrier timed out: a peer rank never arrived" if code == 1 else
                               "sirgcn_peer_push_tma: a bulk copy never landed")

    def ensure_writable(self, sl: PeerSlice):
        """call before the producer overwrites sl.local: fences the peers' pulls of its previous contents"""
        for h in list(sl.pending):
            h.wait()
        # Every rank must issue the SAME number of barriers (the flag epochs are per-rank counters), so the decision
        # may only depend on what the SPMD program did explicitly: `gathered` is set by gather(), never from a
        # finalizer.  (waited_at / joined are written by wait(), which release() may call from _Lease.__del__ at a
        # GC-dependent moment — they are bookkeeping only.)
        if sl.gathered:
            self.barrier(blocking=True)
        sl.gathered = False
        sl.waited_at = -1

    # ---- the all-gather -----------------------------------------------------------------------------------------
    def gather(self, sl: PeerSlice, full, lo=0, hi=None, dst_row_of=None):
        """rows [lo, hi) of every rank's slice -> full[dst_row_of(r) : + hi - lo] (full is [world*rows, ld];
        dst_row_of(r) = r*rows + lo unless the caller lays the table out differently, partition.table_row).  The pulls
        run on the pool's copy streams behind a barrier ("every rank has written these rows") that does not block
        the current stream; the own rows are copied on the current stream.  Returns a PullHandle."""
        hi = sl.rows if hi is None else hi
        if dst_row_of is None:
            dst_row_of = lambda r: r * sl.rows + lo
        if self.world == 1:
            full[dst_row_of(0):dst_row_of(0) + hi - lo].copy_(sl.local[lo:hi])
            return PullHandle(self, sl, None, self.barriers)
        index, done = self.barrier(blocking=False)
        events = []
        nbytes, off = (hi - lo) * sl.row_bytes, lo * sl.row_bytes
        for i in range(1, self.world):
            peer = (self.rank + i) % self.world             # staggered: at any moment the ranks read different peers
            st = self.streams[i - 1]
            st.wait_event(done)
            dst = full.data_ptr() + dst_row_of(peer) * sl.row_bytes
            with torch.cuda.device(self.device):
                rc = self.lib.sirgcn_peer_copy(C.c_void_p(dst), C.c_void_p(sl.peer_ptr[peer] + off), C.c_size_t(nbytes),
                                               C.c_void_p(st.cuda_stream))
            _lib.check(rc, "sirgcn_peer_copy")
            ev = torch.cuda.Event()
            ev.record(st)
            events.append(ev)
        full[dst_row_of(self.rank):dst_row_of(self.rank) + hi - lo].copy_(sl.local[lo:hi])
        h = PullHandle(self, sl, events, index)
        sl.pending.append(h)
        sl.gathered = True
        return h

    # ---- push variant: the gathered tables are the peer-mapped objects ------------------------------------------
    def acquire_full(self, rows, ld, dtype) -> PeerFull:
        """collective the first time a [world*rows, ld] table of this shape is needed; recycled afterwards"""
        key = (rows, ld, dtype)
        lst = self.free_full.setdefault(key, [])
        if lst:
            return lst.pop()
        es = torch.empty((), dtype=dtype).element_size()
        nbytes = max(16, self.world * rows * ld * es)
        base, peers = self._alloc_mapped(nbytes)
        raw = torch.as_tensor(_RawDeviceBuffer(base, nbytes), device=self.device)
        local = raw.view(dtype)[:self.world * rows * ld].view(self.world * rows, ld)
        pf = PeerFull(rows, ld, dtype, base, local, peers)
        self.all_full.append(pf)
        return pf

    def release_full(self, pf: PeerFull):
        """call after the last kernel that reads pf.local has been enqueued on the current stream"""
        pf.released_at = self.barriers
        self.free_full.setdefault((pf.rows, pf.ld, pf.dtype), []).append(pf)

    def push(self, src, pf: PeerFull, lo=0, hi=None, mode="ce", dst_row=None):
        """rows [lo, hi) of this rank's slice `src` ([rows, ld], any local tensor that stays alive until the handle has
        been waited for) -> rows rank*rows + lo.. of EVERY rank's table.  mode "ce": one copy-engine write per peer;
        "sm": one fan-out kernel of a few CTAs (sirgcn_peer_push); "tma": the same with TMA bulk copies issued by one
        thread per CTA (sirgcn_peer_push_tma).  Two barriers frame the transfer, neither blocks
        the current stream: "every rank is done reading the table's previous contents" and "every rank's rows have
        landed".  Returns a PushHandle."""
        hi = pf.rows if hi is None else hi
        cur = torch.cuda.current_stream(self.device)
        # row of the gathered table where src[lo] goes (the same in every rank's table): rank-major unless the caller
        # lays the table out differently (partition.table_row)
        mine = self.rank * pf.rows if dst_row is None else dst_row - lo
        if self.world == 1:
            pf.local[mine + lo:mine + hi].copy_(src[lo:hi])
            return PushHandle(self, None, self.barriers)
        # "every rank is done reading the table's previous contents": always a barrier of its own.  Skipping it when
        # one has been issued since release_full() would make the barrier COUNT depend on when the lease was released
        # — possibly from a finalizer, at a different moment on every rank — and skew the per-rank flag epochs.
        _, free_ev = self.barrier(blocking=False)
        ready = torch.cuda.Event()
        ready.record(cur)                                   # the producer of src[lo:hi]
        nbytes, off = (hi - lo) * pf.row_bytes, (mine + lo) * pf.row_bytes
        src_ptr = src.data_ptr() + lo * src.stride(0) * src.element_size()
        assert src.stride(0) == pf.ld and src.stride(1) == 1
        events = []
        if mode in ("sm", "tma"):
            st = self.push_stream
            st.wait_event(free_ev)
            st.wait_event(ready)
            targets = [pf.peer_ptr[(self.rank + i) % self.world] + off for i in range(1, self.world)]
            targets.append(pf.base + off)
            arr = (C.c_void_p * len(targets))(*targets)
            with torch.cuda.device(self.device):
                if mode == "tma":
                    rc = self.lib.sirgcn_peer_push_tma(C.c_void_p(src_ptr), arr, C.c_int32(len(targets)),
                                                       C.c_size_t(nbytes), C.c_int32(self.push_ctas),
                                                       C.c_void_p(self._status.data_ptr()), C.c_void_p(st.cuda_stream))
                else:
                    rc = self.lib.sirgcn_peer_push(C.c_void_p(src_ptr), arr, C.c_int32(len(targets)),
                                                   C.c_size_t(nbytes), C.c_int32(self.push_ctas),
                                                   C.c_void_p(st.cuda_stream))
            _lib.check(rc, "sirgcn_peer_push")
            ev = torch.cuda.Event()
            ev.record(st)
            events.append(ev)
        else:
            for i in range(1, self.world):
                peer = (self.rank + i) % self.world
                st = self.streams[i - 1]
                st.wait_event(free_ev)
                st.wait_event(ready)
                with torch.cuda.device(self.device):
                    rc = self.lib.sirgcn_peer_copy(C.c_void_p(pf.peer_ptr[peer] + off), C.c_void_p(src_ptr),
                                                   C.c_size_t(nbytes), C.c_void_p(st.cuda_stream))
                _lib.check(rc, "sirgcn_peer_copy")
                ev = torch.cuda.Event()
                ev.record(st)
                events.append(ev)
            pf.local[mine + lo:mine + hi].copy_(src[lo:hi])
        index, done = self.barrier(blocking=False, after=events)
        return PushHandle(self, done, index)

    def close(self):
        torch.cuda.synchronize(self.device)
        for sl in self.all + self.all_full:
            for p in sl.peer_ptr:
                if p:
                    self.lib.sirgcn_peer_close(C.c_void_p(p))
        for p in self._pad_peers:
            if p:
                self.lib.sirgcn_peer_close(C.c_void_p(p))
        if self.world > 1:
            dist.barrier(group=self.group)      # nobody frees memory a peer still has mapped
        for sl in self.all + self.all_full:
            sl.local = None
            self.lib.sirgcn_peer_free(C.c_void_p(sl.base))
        self.lib.sirgcn_peer_free(C.c_void_p(self._pad_base))
        self.all, self.free, self.all_full, self.free_full = [], {}, [], {}
