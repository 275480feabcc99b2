"""1-D destination-row partition of one large graph over the GPUs of a node (one process per GPU,
torch.distributed / NCCL + peer memory over NVLink).  New capability — the reference is single-process — whose
correctness oracle is "G-rank result == 1-rank result" (SURVEY.md §8e).

Rank r owns node rows [r*n_pad, min(N, (r+1)*n_pad)), n_pad = ceil(N/G): its slice of H, Q, K, A, the
in-CSR rows of those destinations (column ids stay GLOBAL source ids) and the out-CSC rows of those
sources (row ids stay GLOBAL destination ids).  Per layer three row tables travel (each rank receives the other
ranks' slices); the schedule keeps the NVLink ports busy while the edge walks run and exposes only the first one:

    forward   K_1      gathered before the first walk                                         (exposed)
              K_l+1    produced and gathered in destination chunks WHILE layer l's walk runs: chunk c of
                       out_l = A_l[c]·W_R^T -> K_l+1[c] = out_l[c]·W_K^T -> pulls of chunk c travel under the walk of c+1
              Q_L      (last layer) travels behind its own forward walk; it is only needed by the backward CSC pass
    backward  dA_l    *= dst coefficient, then travels while  dQ_l = edge_bwd_q over the local CSR rows (K_l) runs
              Q_l-1    travels while  dK_l = edge_bwd_k over the local CSC rows (Q_l, dA_l, local K_l) runs
    weights   replicated; all layers' dW summed by ONE flat all_reduce at the end of the backward pass

Gathered tables are laid out [G*n_pad, ld], so a global node id indexes them directly.  How tables travel is the
transport: "peer" = copy-engine pulls from IPC-mapped peer slices (peer.py; takes no SM from the walks), or
"collective" = torch.distributed all-gathers (NCCL kernels; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib, gemm
from . import function as F_
from .graph import DEFAULT_LONG_THRESHOLD, CompressedRows, Graph


class CudaEdgeBackend:
    """the product backend: C-ABI CUDA kernels"""
    forward = staticmethod(F_.edge_forward)
    backward_q = staticmethod(F_.edge_backward_q)
    backward_k = staticmethod(F_.edge_backward_k)


def build_rows(key, other, n_rows, long_threshold=DEFAULT_LONG_THRESHOLD):
    """CompressedRows over rows = `key` values in [0, n_rows) carrying `other` (any int32 payload, e.g. global
    node ids), by the stable GPU sort of sirgcn_rows_build."""
    if not key.is_cuda:
        raise RuntimeError("build_rows needs CUDA index tensors (no CPU fallback)")
    dev = key.device
    key, other = key.to(torch.int32).contiguous(), other.to(torch.int32).contiguous()
    e = int(key.numel())
    indptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
    out = torch.empty(e, dtype=torch.int32, device=dev)
    L = _lib.lib()
    nbytes = L.sirgcn_rows_build_workspace_bytes(C.c_int64(e), C.c_int32(n_rows))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.sirgcn_rows_build(_lib.ptr(key), _lib.ptr(other), C.c_int64(e), C.c_int32(n_rows), _lib.ptr(indptr),
                                 _lib.ptr(out), None, _lib.ptr(ws), C.c_size_t(nbytes), _lib.stream_ptr(dev))
    _lib.check(rc, "sirgcn_rows_build")
    del ws
    return CompressedRows(indptr, out, None, long_threshold)


class _CollectiveSlice:
    def __init__(self, local):
        self.local = local


class _Handle:
    """completion of one (chunk of a) table gather: wait() makes the CURRENT STREAM wait, never the host"""

    def __init__(self, works):
        self.works = works

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []


class CollectiveTransport:
    """all-gathers through torch.distributed (NCCL kernels on the GPU, gloo in the CPU tests of the host logic)"""
    kind = "collective"

    def __init__(self, part):
        self.part = part

    def acquire(self, rows, ld, dtype, device, zero):
        return _CollectiveSlice((torch.zeros if zero else torch.empty)((rows, ld), dtype=dtype, device=device))

    def new_full(self, sl, lease):
        return sl.local.new_empty((self.part.world * sl.local.shape[0], sl.local.shape[1]))

    def gather(self, sl, full, lo=0, hi=None):
        """rows [lo, hi) of every rank's slice -> their rows of the gathered table (RowPartition.table_row); returns a
        handle.  A whole layout chunk of every rank lands in ONE contiguous block of the table, so it travels by
        all_gather_into_tensor straight into place — no list of views, no staging copy."""
        part, rows = self.part, sl.local.shape[0]
        hi = rows if hi is None else hi
        G, step = part.world, part.step
        if G == 1:
            full[lo:hi].copy_(sl.local[lo:hi])
            return _Handle([])
        works = []
        for a, b, c in part.chunk_pieces(lo, hi):
            if a == c * step and b == (c + 1) * step:
                works.append(dist.all_gather_into_tensor(full[c * G * step:(c + 1) * G * step], sl.local[a:b],
                                                         group=part.group, async_op=True))
            else:       # part of a layout chunk: the ranks' pieces are not adjacent in the table
                views = [full[part.table_row(r, a):part.table_row(r, a) + (b - a)] for r in range(G)]
                works.append(dist.all_gather(views, sl.local[a:b], group=part.group, async_op=True))
        return _Handle(works)

    def release(self, sl):
        pass


class PeerTransport:
    """all-gathers as copy-engine pulls from IPC-mapped peer slices (peer.py): no SM taken from the edge walks"""
    kind = "peer"

    def __init__(self, part):
        from . import peer
        self.part = part
        self.pool = peer.PeerPool(part.group)

    def acquire(self, rows, ld, dtype, device, zero):
        sl = self.pool.acquire(rows, ld, dtype)
        self.pool.ensure_writable(sl)
        if zero:
            sl.local.zero_()
        return sl

    def new_full(self, sl, lease):
        return sl.local.new_empty((self.pool.world * sl.rows, sl.ld))

    def gather(self, sl, full, lo=0, hi=None):
        part = self.part
        hs = [self.pool.gather(sl, full, a, b, dst_row_of=lambda r, a=a: part.table_row(r, a))
              for a, b, _ in part.chunk_pieces(lo, sl.rows if hi is None else hi)]
        return hs[0] if len(hs) == 1 else _Handle(hs)

    def release(self, sl):
        self.pool.release(sl)


class _FullToken:
    def __init__(self, pf):
        self.pf = pf


class PushTransport:
    """all-gathers as PUSHES into IPC-mapped gathered tables (peer.py): every rank writes its slice into every
    peer's table — by copy engines ("push") or by a fan-out kernel of a few CTAs ("pushsm")"""

    def __init__(self, part, mode):
        from . import peer
        self.part = part
        self.pool = peer.PeerPool(part.group)
        self.mode = mode
        self.kind = {"ce": "push", "sm": "pushsm", "tma": "pushtma"}[mode]
        self._fulls = {}

    def acquire(self, rows, ld, dtype, device, zero):
        return _CollectiveSlice((torch.zeros if zero else torch.empty)((rows, ld), dtype=dtype, device=device))

    def new_full(self, sl, lease):
        pf = self.pool.acquire_full(sl.local.shape[0], sl.local.shape[1], sl.local.dtype)
        self._fulls[pf.local.data_ptr()] = pf
        lease.add(_FullToken(pf))
        return pf.local

    def gather(self, sl, full, lo=0, hi=None):
        part, pf = self.part, self._fulls[full.data_ptr()]
        hs = [self.pool.push(sl.local, pf, a, b, self.mode, dst_row=part.table_row(part.rank, a))
              for a, b, _ in part.chunk_pieces(lo, pf.rows if hi is None else hi)]
        return hs[0] if len(hs) == 1 else _Handle(hs)

    def release(self, obj):
        if isinstance(obj, _FullToken):
            self.pool.release_full(obj.pf)


class _Lease:
    """slices held by one forward call until its backward has run (or its autograd node is dropped)"""

    def __init__(self, transport):
        self.transport, self.slices = transport, []

    def add(self, sl):
        self.slices.append(sl)
        return sl

    def release(self):
        for sl in self.slices:
            self.transport.release(sl)
        self.slices = []

    def __del__(self):
        try:
            self.release()
        except Exception:       # interpreter shutdown
            pass


class RowPartition:
    """Local slice of a graph for rank `rank` of `world` (see module docstring).

    csr: rows = local destinations, idx = sources;  csc: rows = local sources, idx = destinations — idx values are
    ROWS OF THE GATHERED TABLES (table_row of the global node id), handed in as global ids and remapped here.
    in_norm / out_norm / inv_in_deg: fp32 [G*n_pad] coefficient vectors in gathered-table order (padding = 1).

    Gathered-table layout (`layout_chunks` = C): every rank's padded slice [n_pad = C*step rows] is cut into C chunks;
    the table holds chunk 0 of rank 0..G-1, then chunk 1 of rank 0..G-1, ... — row (c*G + r)*step + j for local row
    c*step + j of rank r.  Chunk c of ALL ranks is therefore one contiguous block that an all_gather_into_tensor
    writes in place, which is what lets the cross-layer prefetch send a table chunk by chunk without staging copies
    (C = 1 is the plain rank-major table, row = global node id)."""

    def __init__(self, num_nodes, rank, world, csr_local, csc_local, in_norm, out_norm, inv_in_deg, group=None,
                 transport="auto", layout_chunks=1):
        self.num_nodes_, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        self._transport, self._transport_kind = None, transport
        self.layout_chunks = max(1, int(layout_chunks))
        self.n_pad, self.lo, self.hi = self.bounds(self.num_nodes_, rank, world, self.layout_chunks)
        self.step = self.n_pad // self.layout_chunks
        self.csr, self.csc = csr_local, csc_local
        n1 = max(self.hi - self.lo, 1)
        # coefficients of the LOCAL rows (row side of the walks), taken before the table order is applied
        self._local = {"in_norm": in_norm[self.lo:self.lo + n1].clone(), "out_norm": out_norm[self.lo:self.lo + n1].clone(),
                       "inv_in_deg": inv_in_deg[self.lo:self.lo + n1].clone()}
        if self.layout_chunks > 1 and world > 1:
            reorder = lambda v: v.view(world, self.layout_chunks, self.step).transpose(0, 1).reshape(-1).contiguous()
            in_norm, out_norm, inv_in_deg = reorder(in_norm), reorder(out_norm), reorder(inv_in_deg)
            self._remap_(self.csr.idx)
            self._remap_(self.csc.idx)
        self.in_norm, self.out_norm, self.inv_in_deg = in_norm, out_norm, inv_in_deg
        self.num_local_edges = csr_local.num_pos
        self._row_chunks = {}

    @staticmethod
    def bounds(num_nodes, rank, world, layout_chunks=1):
        """(n_pad, lo, hi): rows per padded slice (a multiple of layout_chunks) and this rank's global row range"""
        c = max(1, int(layout_chunks))
        n_pad = ((num_nodes + world - 1) // world + c - 1) // c * c
        lo = min(num_nodes, rank * n_pad)
        return n_pad, lo, min(num_nodes, lo + n_pad)

    # ---- gathered-table layout -------------------------------------------------------------------------------
    def table_row(self, r, i):
        """row of the gathered tables that holds local row i of rank r"""
        c = i // self.step
        return (c * self.world + r) * self.step + (i - c * self.step)

    def chunk_pieces(self, lo, hi):
        """[(a, b, c)]: the local row range [lo, hi) cut at layout-chunk boundaries (piece [a, b) lies in chunk c)"""
        out, a = [], lo
        while a < hi:
            c = a // self.step
            b = min(hi, (c + 1) * self.step)
            out.append((a, b, c))
            a = b
        return out

    def _remap_(self, idx, block=1 << 26):
        """global node ids -> gathered-table rows, in place (bounded temporaries: the 2 B-edge graph has 250 M per rank)"""
        G, n_pad, step = self.world, self.n_pad, self.step
        for a in range(0, idx.numel(), block):
            g = idx[a:a + block]
            r = torch.div(g, n_pad, rounding_mode="floor")
            i = g - r * n_pad
            c = torch.div(i, step, rounding_mode="floor")
            g.copy_((c * G + r) * step + (i - c * step))

    @property
    def n_local(self):
        return self.hi - self.lo

    def row_chunks(self, chunks):
        """[(lo, hi, CompressedRows or None)]: the local destination rows cut into `chunks` equal ranges of the PADDED
        slice (the same bounds on every rank, so chunk c of every rank's K slice travels together); the structure
        covers rows [lo, min(hi, n_local)) and is None when that range is empty."""
        got = self._row_chunks.get(chunks)
        if got is None:
            step = self._cut_step(chunks)
            got = []
            for c in range(chunks):
                lo, hi = min(self.n_pad, c * step), min(self.n_pad, (c + 1) * step)
                top = min(hi, self.n_local)
                rows = self.csr if (chunks == 1 and top == self.n_local) else \
                    (self.csr.slice_rows(lo, top) if top > lo else None)
                got.append((lo, hi, rows))
            self._row_chunks[chunks] = got
        return got

    def _cut_step(self, chunks):
        """rows per cut of the walks: whole layout chunks when the table is chunk-major (cuts then coincide with the
        contiguous blocks of the gathered tables)"""
        if self.layout_chunks > 1 and self.layout_chunks % chunks == 0:
            return self.step * (self.layout_chunks // chunks)
        return (self.n_pad + chunks - 1) // chunks

    def col_chunks(self, chunks):
        """the same cut of the local SOURCE rows (out-CSC), for the chunked dK walk of the backward pass"""
        got = self._row_chunks.get(("csc", chunks))
        if got is None:
            step = self._cut_step(chunks)
            got = []
            for c in range(chunks):
                lo, hi = min(self.n_pad, c * step), min(self.n_pad, (c + 1) * step)
                top = min(hi, self.n_local)
                rows = self.csc if (chunks == 1 and top == self.n_local) else \
                    (self.csc.slice_rows(lo, top) if top > lo else None)
                got.append((lo, hi, rows))
            self._row_chunks[("csc", chunks)] = got
        return got

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_csr_csc(cls, csr: CompressedRows, csc: CompressedRows, num_nodes, rank, world, group=None,
                     in_norm=None, out_norm=None, inv_in_deg=None, transport="auto", layout_chunks=1):
        """slice replicated whole-graph structures (tests, small graphs)"""
        n = int(num_nodes)
        n_pad, lo, hi = cls.bounds(n, rank, world, layout_chunks)
        if in_norm is None:
            in_deg = (csr.indptr[1:] - csr.indptr[:-1]).clamp(min=1).to(torch.float32)
            out_deg = (csc.indptr[1:] - csc.indptr[:-1]).clamp(min=1).to(torch.float32)
            in_norm, out_norm, inv_in_deg = 1.0 / torch.sqrt(in_deg), 1.0 / torch.sqrt(out_deg), 1.0 / in_deg
        pad = lambda t: torch.cat([t, t.new_ones(world * n_pad - n)]) if world * n_pad > n else t
        return cls(n, rank, world, csr.slice_rows(lo, hi), csc.slice_rows(lo, hi),
                   pad(in_norm), pad(out_norm), pad(inv_in_deg), group, transport, layout_chunks)

    @classmethod
    def from_graph(cls, graph: Graph, rank, world, group=None, transport="auto", layout_chunks=1):
        """slice an already converted (replicated) Graph; the caller may drop `graph` afterwards"""
        return cls.from_csr_csc(graph.csr, graph.csc, graph.num_nodes(), rank, world, group,
                                graph.in_norm, graph.out_norm, graph.inv_in_deg, transport, layout_chunks)

    @classmethod
    def from_local_edges(cls, num_nodes, rank, world, in_src, in_dst, out_src, out_dst, group=None,
                         long_threshold=DEFAULT_LONG_THRESHOLD, in_indptr=None, transport="auto", layout_chunks=1):
        """build from this rank's two edge lists (GLOBAL ids): the edges whose destination is local
        (in_src -> in_dst) and the edges whose source is local (out_src -> out_dst).  When the first list is
        already destination-sorted, pass its local row pointer `in_indptr` instead of `in_dst`.  Degree
        coefficients of the whole graph are assembled with one all_gather each."""
        n = int(num_nodes)
        n_pad, lo, hi = cls.bounds(n, rank, world, layout_chunks)
        dev = in_src.device
        nl = hi - lo
        if in_indptr is not None:
            csr = CompressedRows(in_indptr.to(torch.int32).contiguous(), in_src.to(torch.int32).contiguous(), None,
                                 long_threshold)
        else:
            csr = build_rows(in_dst - lo, in_src, nl, long_threshold)
        csc = build_rows(out_src - lo, out_dst, nl, long_threshold)
        in_deg = (csr.indptr[1:] - csr.indptr[:-1]).clamp(min=1).to(torch.float32)
        out_deg = (csc.indptr[1:] - csc.indptr[:-1]).clamp(min=1).to(torch.float32)

        def gather(v):
            buf = torch.ones(n_pad, dtype=torch.float32, device=dev)
            buf[:nl] = v
            full = torch.empty(world * n_pad, dtype=torch.float32, device=dev)
            if world > 1:
                dist.all_gather_into_tensor(full, buf, group=group)
            else:
                full.copy_(buf)
            return full
        return cls(n, rank, world, csr, csc, gather(1.0 / torch.sqrt(in_deg)), gather(1.0 / torch.sqrt(out_deg)),
                   gather(1.0 / in_deg), group, transport, layout_chunks)

    @classmethod
    def synthetic_powerlaw(cls, num_nodes, num_edges, rank, world, alpha=2.3, max_deg=None, seed=0, device="cuda",
                           group=None, long_threshold=DEFAULT_LONG_THRESHOLD, transport="auto", layout_chunks=1):
        """this rank's slice of synth.powerlaw_hashed (the SAME global graph on every world size), generated on
        the device without any edge exchange"""
        from . import synth
        indptr = synth.powerlaw_indptr(num_nodes, num_edges, alpha, max_deg, seed, device)
        n_pad, lo, hi = cls.bounds(num_nodes, rank, world, layout_chunks)
        in_src, _ = synth.powerlaw_hashed_rows(indptr, num_nodes, lo, hi, seed, want_dst=False)
        out_src, out_dst = synth.powerlaw_hashed_cols(indptr, num_nodes, lo, hi, seed)
        in_indptr = indptr[lo:hi + 1] - indptr[lo]
        del indptr
        return cls.from_local_edges(num_nodes, rank, world, in_src, None, out_src, out_dst, group, long_threshold,
                                    in_indptr=in_indptr, transport=transport, layout_chunks=layout_chunks)

    # ---- coefficients --------------------------------------------------------------------------------------
    def scales_rows(self, agg_type):
        """(dst_scale, src_scale) for the CSR walks: dst = local row, src = row of the gathered K table"""
        if agg_type == "sym":
            return self._local["in_norm"], self.out_norm
        if agg_type == "mean":
            return self._local["inv_in_deg"], None
        return None, None

    def scale_cols_rows(self, agg_type):
        """row (= local source) scale of the CSC walk; its destination scale is folded into dA before the gather"""
        return self._local["out_norm"] if agg_type == "sym" else None

    # ---- transport --------------------------------------------------------------------------------------------
    def transport(self):
        """how row tables travel between the ranks (CUDA tensors of an NCCL group of one node, except "collective"):
        "push" / "pushsm" / "pushtma" (every rank writes its slice into every peer's IPC-mapped gathered table, by
        copy engines / a fan-out kernel / a TMA bulk-copy fan-out kernel), "peer" (copy-engine pulls from IPC-mapped peer slices) or "collective"
        (torch.distributed all-gathers).  "auto" picks auto_transport(world) when it applies; the environment variable
        SIRGCN_TRANSPORT overrides."""
        if self._transport is None:
            import os
            kind = os.environ.get("SIRGCN_TRANSPORT", self._transport_kind)
            if kind == "auto":
                on_gpu = self.csr.indptr.is_cuda and self.world > 1 and dist.is_initialized() \
                    and dist.get_backend(self.group) == "nccl"
                kind = auto_transport(self.world) if on_gpu else "collective"
            make = {"peer": lambda: PeerTransport(self), "push": lambda: PushTransport(self, "ce"),
                    "pushsm": lambda: PushTransport(self, "sm"), "pushtma": lambda: PushTransport(self, "tma"),
                    "collective": lambda: CollectiveTransport(self)}[kind]
            requested = os.environ.get("SIRGCN_TRANSPORT", self._transport_kind)
            try:
                self._transport = make()
            except RuntimeError as exc:     # peer-memory setup failed on some rank (agreed on by all ranks, peer.py)
                if requested != "auto":
                    raise
                import warnings
                warnings.warn(f"peer-memory transport unavailable ({exc}); using torch.distributed all-gathers")
                self._transport = CollectiveTransport(self)
        return self._transport

    def all_gather_rows(self, local, out=None):
        """[n_local, c] rows of this rank -> [G*n_pad, c] with rank r's rows at r*n_pad (padding rows are zero).
        For INPUT tables such as the node features: gathered once per step ahead of the layers (a prefetching
        loader does it under the previous step), they let layer 1 project its K and Q tables locally instead of
        gathering both on the critical path (`feat_full` of partitioned_sirconv_stack)."""
        full = local.new_empty((self.world * self.n_pad, local.shape[1])) if out is None else out
        src = local
        if local.shape[0] != self.n_pad:
            src = local.new_zeros((self.n_pad, local.shape[1]))
            src[:local.shape[0]].copy_(local)
        if self.world == 1:
            full.copy_(src)
        elif self.layout_chunks == 1:
            dist.all_gather_into_tensor(full, src.contiguous(), group=self.group)
        else:           # gathered-table order: chunk c of every rank is one contiguous block
            src, G, step = src.contiguous(), self.world, self.step
            for c in range(self.layout_chunks):
                dist.all_gather_into_tensor(full[c * G * step:(c + 1) * G * step], src[c * step:(c + 1) * step],
                                            group=self.group)
        return full

    def all_reduce_(self, t):
        if self.world > 1 and t is not None:
            dist.all_reduce(t, group=self.group)
        return t


def auto_transport(world):
    """NCCL all-gathers at every world size: the best schedule MEASURED inside the step (8 GPUs, 2 B-edge graph:
    160 ms with NCCL against 207 ms with copy-engine pushes — a push that takes 21 ms alone takes ~48 ms under an
    HBM-saturating walk, SCALE_r01.json).  The peer-memory transports stay selectable (`transport=` /
    SIRGCN_TRANSPORT); a default only changes on an in-step A/B committed under profiles/."""
    return "collective"


# bench.py sets this to a list to collect (label, CUDA event) marks on the compute stream (phase breakdown)
PHASE_MARKS = None


def _mark(label):
    if PHASE_MARKS is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PHASE_MARKS.append((label, ev))


def _project(x, w, b, dst, n, d, ld):
    """dst[:n, :d] = x·w^T + b (dst is a [*, ld] slice buffer)"""
    if n == 0:
        return
    if ld == d:
        gemm.linear_forward(x, w, b, out=dst[:n])
    else:
        dst[:n, :d].copy_(gemm.linear_forward(x, w, b))


def _table(buf, n, d):
    t = buf[:n, :d]
    t._sirgcn_padded = True
    return t


class _walk_mode:
    """plain-grid edge walks while a stack runs on more than one rank (the collectives' kernels must be able to become
    resident under a walk; persistent walk CTAs would hold every SM until the walk ends — function.WALK_PERSISTENT).
    SIRGCN_PARTITION_PERSIST=1 keeps the persistent walks (for measurements)."""

    def __init__(self, part):
        import os
        self.plain = part.world > 1 and os.environ.get("SIRGCN_PARTITION_PERSIST", "0") != "1"

    def __enter__(self):
        self.saved = F_.WALK_PERSISTENT
        if self.plain:
            F_.WALK_PERSISTENT = False

    def __exit__(self, *exc):
        F_.WALK_PERSISTENT = self.saved
        return False


class PartitionedSIRStackFunction(torch.autograd.Function):
    """L stacked SIRConv layers (sum / mean / sym, elementwise σ, no dropout, nothing between the layers) on a
    row-partitioned graph as ONE autograd node — the unit that can hide the table traffic (module docstring):
    same arithmetic per row as SIRLayerFunction; every rank ends with the full-graph weight gradients.

    apply(feat, part, cfgs, opts, backend, feat_full, *weights):  cfgs[l] = (agg_type, act, act_param);
    weights = (w_q, b_q, w_k, w_r, b_r) per layer, flattened; opts = {"chunks": destination chunks of the cross-layer
    prefetch, "gather": "inputs" | "projections"}.
    "inputs": what travels for layer l+1 is its INPUT H_l+1 (= layer l's output rows, sent chunk by chunk under layer
    l's walk); every rank then projects the whole K table (forward) and Q table (backward) of layer l+1 locally —
    one table transfer per layer instead of two, for two 50 M-row GEMMs (4 ms each against a 17-19 ms transfer).
    "projections": K_l+1 is projected per chunk and travels under layer l's walk, Q_l travels in backward.
    feat_full (optional, no gradient): the input rows of ALL ranks (RowPartition.all_gather_rows); the first layer
    then projects its tables locally in either mode."""

    @staticmethod
    def forward(ctx, feat, part: RowPartition, cfgs, opts, backend, feat_full, *weights):
        with _walk_mode(part):
            return PartitionedSIRStackFunction._forward(ctx, feat, part, cfgs, opts, backend, feat_full, *weights)

    @staticmethod
    def _forward(ctx, feat, part: RowPartition, cfgs, opts, backend, feat_full, *weights):
        L = len(cfgs)
        W = [weights[5 * l:5 * l + 5] for l in range(L)]
        n, dt, dev = part.n_local, feat.dtype, feat.device
        tr = part.transport()
        leases = [_Lease(tr) for _ in range(L)]     # what layer l holds until its backward has run
        train = any(ctx.needs_input_grad)
        chunks = max(1, int(opts.get("chunks", 4))) if part.world > 1 else 1
        gather_inputs = opts.get("gather", "projections") == "inputs"
        state = []
        h = feat
        pre = None                          # (k_sl, k_all, [handles]) of the current layer when K was prefetched
        h_full, h_handles = feat_full, []   # this layer's input rows of ALL ranks (+ their arrival), when they travel
        for l in range(L):
            lease = leases[l]
            w_q, b_q, w_k, w_r, b_r = W[l]
            agg_type, act, act_param = cfgs[l]
            d = w_q.shape[0]
            ld = F_._pad_cols(d, dt)
            _mark("fwd:start")
            if h_full is not None:
                for hnd in h_handles:
                    hnd.wait()
                _mark("fwd:wait_H")
                rows_all = h_full.shape[0]
                keep = opts.get("keep_q_full")
                if keep is None:    # auto: ONE [Q|K] GEMM over the gathered input, Q kept for the backward CSC walk,
                    # when the extra table fits comfortably (from 4 GPUs up; at 2 GPUs the 2 B-edge graph leaves no room for it)
                    keep = train and (not h_full.is_cuda or (part.world >= 4 and
                                      torch.cuda.mem_get_info(dev)[0] > 4 * rows_all * ld * h_full.element_size()))
                q_all_kept = None
                if keep and train:
                    w_cat_f = (torch.zeros if ld != d else torch.empty)((2 * ld, w_q.shape[1]), dtype=dt, device=dev)
                    w_cat_f[:d].copy_(w_q)
                    w_cat_f[ld:ld + d].copy_(w_k)
                    b_cat_f = None
                    if b_q is not None:
                        # in the parameter's own dtype: the projection adds the bias in fp32, exactly as the local Q
                        # projection below does — a bias rounded to the table dtype here would give the backward CSC
                        # walk (which reads THIS table) other pre-activation signs than the forward walk saw
                        b_cat_f = torch.zeros(2 * ld, dtype=b_q.dtype, device=dev)
                        b_cat_f[:d].copy_(b_q)
                    qk_all = gemm.linear_forward(h_full.to(dt), w_cat_f, b_cat_f)       # [rows_all, 2·ld]
                    k_all = qk_all[:, ld:2 * ld]            # row stride 2·ld: the walks take any 16-byte row pitch
                    q_all_kept = qk_all[:, :ld]
                    del qk_all
                else:
                    k_all = torch.empty((rows_all, ld), dtype=dt, device=dev) if ld == d else \
                        torch.zeros((rows_all, ld), dtype=dt, device=dev)
                    _project(h_full.to(dt), w_k.to(dt), None, k_all, rows_all, d, ld)
                # this rank's own K rows (row operand and output of the backward CSC walk): projected from the local
                # input — in a chunk-major table they are not one contiguous slice of k_all
                k_sl = _CollectiveSlice((torch.zeros if ld != d else torch.empty)((part.n_pad, ld), dtype=dt, device=dev))
                if train:
                    _project(h, w_k.to(dt), None, k_sl.local, n, d, ld)
                k_handles = []
            elif pre is None:
                k_sl = lease.add(tr.acquire(part.n_pad, ld, dt, dev, zero=ld != d))
                _project(h, w_k.to(dt), None, k_sl.local, n, d, ld)
                k_all = tr.new_full(k_sl, lease)
                k_handles = [tr.gather(k_sl, k_all)]
            else:
                k_sl, k_all, k_handles = pre
                pre = None
            q_sl = lease.add(tr.acquire(part.n_pad, ld, dt, dev, zero=ld != d))
            _project(h, w_q.to(dt), b_q, q_sl.local, n, d, ld)
            # Q of the remote destinations is consumed by the backward CSC pass only.  The last layer's travels behind
            # its own forward walk; an earlier layer's is left for the backward pass (the ports carry K_l+1 now).
            q_full = q_h = None
            if train and l == L - 1 and h_full is None:
                q_full = tr.new_full(q_sl, lease)
                q_h = tr.gather(q_sl, q_full)
            q, kf = _table(q_sl.local, n, d), _table(k_all, k_all.shape[0], d)
            ds, ss = part.scales_rows(agg_type)
            _mark("fwd:proj")
            a = F_._alloc_table(n, d, dt, dev, zero=ld != d)
            for hnd in k_handles:
                hnd.wait()
            _mark("fwd:wait_K")
            h_full_next, h_handles_next = None, []
            if l + 1 < L and gather_inputs:
                # walk in destination chunks; chunk c of this layer's OUTPUT (the next layer's input) is sent to every
                # rank while chunk c+1 is walked
                do = w_r.shape[0]
                ldo = F_._pad_cols(do, dt)
                out_sl = leases[l + 1].add(tr.acquire(part.n_pad, ldo, dt, dev, zero=ldo != do))
                hn_all = tr.new_full(out_sl, leases[l + 1])
                for lo, hi, rows in part.row_chunks(chunks):
                    if rows is not None:
                        top = lo + rows.n_rows
                        backend.forward(rows, q[lo:top], kf, None, None if ds is None else ds[lo:top], ss, act,
                                        act_param, out=a[lo:top])
                        if ldo == do:
                            gemm.linear_forward(a[lo:top], w_r.to(dt), b_r, out=out_sl.local[lo:top])
                        else:
                            out_sl.local[lo:top, :do].copy_(gemm.linear_forward(a[lo:top], w_r.to(dt), b_r))
                    if hi > lo:
                        h_handles_next.append(tr.gather(out_sl, hn_all, lo, hi))
                out = out_sl.local[:n, :do]
                h_full_next = hn_all[:, :do]
                a._sirgcn_padded = True
                _mark("fwd:edge")
            elif l + 1 < L:
                # walk in destination chunks; chunk c of the NEXT layer's K is produced and sent while c+1 is walked
                w_kn = W[l + 1][2].to(dt)
                dn = w_kn.shape[0]
                ldn = F_._pad_cols(dn, dt)
                kn_sl = leases[l + 1].add(tr.acquire(part.n_pad, ldn, dt, dev, zero=ldn != dn))
                kn_all = tr.new_full(kn_sl, leases[l + 1])
                out = torch.empty((n, w_r.shape[0]), dtype=dt, device=dev)
                kn_handles = []
                for lo, hi, rows in part.row_chunks(chunks):
                    if rows is not None:
                        top = lo + rows.n_rows
                        backend.forward(rows, q[lo:top], kf, None, None if ds is None else ds[lo:top], ss, act,
                                        act_param, out=a[lo:top])
                        gemm.linear_forward(a[lo:top], w_r.to(dt), b_r, out=out[lo:top])
                        if ldn == dn:
                            gemm.linear_forward(out[lo:top], w_kn, None, out=kn_sl.local[lo:top])
                        else:
                            kn_sl.local[lo:top, :dn].copy_(gemm.linear_forward(out[lo:top], w_kn, None))
                    if hi > lo:
                        kn_handles.append(tr.gather(kn_sl, kn_all, lo, hi))
                pre = (kn_sl, kn_all, kn_handles)
                a._sirgcn_padded = True
                _mark("fwd:edge")
            else:
                backend.forward(part.csr, q, kf, None, ds, ss, act, act_param, out=a)
                a._sirgcn_padded = True
                _mark("fwd:edge")
                out = gemm.linear_forward(a, w_r.to(dt), b_r)
            _mark("fwd:out")
            # intermediates live on ctx (not in save_for_backward) so that backward can drop layer l's tables as soon
            # as layer l is done — the whole stack is ONE autograd node, whose saved tensors would otherwise all live
            # until its backward returns (at 2 GPUs that is the difference between fitting the 2 B-edge graph or not)
            state.append(dict(k_sl=k_sl, q_sl=q_sl, q_full=q_full, q_h=q_h, d=d, ld=ld, k_all=k_all, a=a,
                              h=None if l == 0 else h, h_full=None if l == 0 else h_full, has_full=h_full is not None,
                              q_all_kept=q_all_kept if h_full is not None else None))
            q_all_kept = None
            h, h_full, h_handles = out, h_full_next, h_handles_next
        if not train:
            for lease in leases:
                lease.release()
            return h
        ctx.save_for_backward(feat, feat_full, *weights)
        ctx.leases, ctx.state, ctx.part, ctx.cfgs, ctx.backend, ctx.L = leases, state, part, cfgs, backend, L
        ctx.bwd_chunks = max(1, int(opts.get("bwd_chunks", 1))) if part.world > 1 else 1
        return h

    @staticmethod
    def backward(ctx, gout):
        with _walk_mode(ctx.part):
            return PartitionedSIRStackFunction._backward(ctx, gout)

    @staticmethod
    def _backward(ctx, gout):
        L, part, be, leases, state = ctx.L, ctx.part, ctx.backend, ctx.leases, ctx.state
        tensors = ctx.saved_tensors
        feat0, feat_full, weights = tensors[0], tensors[1], tensors[2:]
        tr = part.transport()
        n = part.n_local
        need = ctx.needs_input_grad
        wneed = [need[6 + 5 * l:11 + 5 * l] for l in range(L)]
        grads = [[None] * 5 for _ in range(L)]       # fp32 partials of this rank, reduced at the end
        g = gout
        prepared = None                 # (da_sl, da_full, [handles]) of the current layer when the layer above sent it
        for l in reversed(range(L)):
            st, lease = state[l], leases[l]
            feat, k_all, a = (feat0 if l == 0 else st["h"]), st["k_all"], st["a"]
            h_full = (feat_full if l == 0 else st["h_full"]) if st["has_full"] else None
            w_q, b_q, w_k, w_r, b_r = weights[5 * l:5 * l + 5]
            agg_type, act, act_param = ctx.cfgs[l]
            d, ld, q_sl, k_sl = st["d"], st["ld"], st["q_sl"], st["k_sl"]
            dt, dev = q_sl.local.dtype, q_sl.local.device
            adt = torch.float64 if dt == torch.float64 else torch.float32
            g = g.to(dt)
            g = g if g.stride(-1) == 1 else g.contiguous()
            _mark("bwd:start")
            ds, ss = part.scales_rows(agg_type)
            if prepared is None:
                # dA, scaled by the destination coefficient BEFORE it travels: the CSC pass then needs no scale lookup
                da_sl = lease.add(tr.acquire(part.n_pad, ld, dt, dev, zero=ld != d))
                da = _table(da_sl.local, n, d)
                if n:
                    if ld == d:
                        gemm.linear_dgrad(g, w_r.to(dt), out=da)
                    else:
                        da.copy_(gemm.linear_dgrad(g, w_r.to(dt)))
                if ds is not None:
                    da.mul_(ds[:n].to(dt).unsqueeze(1))
                da_full = tr.new_full(da_sl, lease)
                da_handles = [tr.gather(da_sl, da_full)]                  # travels while dQ is computed
            else:
                da_sl, da_full, da_handles = prepared                     # sent chunk by chunk under the dK walk above
                prepared = None
                da = _table(da_sl.local, n, d)
            if wneed[l][3]:     # dW_R and db_R in one pass over g (tcgen05 wgrad kernel for 16-bit tables)
                grads[l][3], db_ = gemm.linear_wgrad_bias(g, a, adt, bool(wneed[l][4] and b_r is not None))
                if db_ is not None:
                    grads[l][4] = db_
            elif wneed[l][4] and b_r is not None:
                grads[l][4] = gemm.column_sum(g, adt)
            q, k = _table(q_sl.local, n, d), _table(k_sl.local, n, d)
            kf, daf = _table(k_all, k_all.shape[0], d), _table(da_full, da_full.shape[0], d)
            # dQ and dK are the two halves of ONE [n, 2·ld] buffer: the projection's weight and input gradients are
            # then single GEMMs over the concatenated [W_Q; W_K]
            dqk = (torch.empty if ld == d else torch.zeros)((n, 2 * ld), dtype=dt, device=dev)
            dq, dk = _table(dqk, n, d), dqk[:, ld:ld + d]
            dk._sirgcn_padded = True
            _mark("bwd:dA")
            be.backward_q(part.csr, q, kf, None, da, None, ss, act, act_param, False, out=dq)
            _mark("bwd:edge_q")
            if h_full is not None and st["q_all_kept"] is not None:
                st["q_full"] = st["q_all_kept"]             # made by the forward pass's one [Q|K] GEMM over the input
                st["q_all_kept"] = None
            elif h_full is not None:
                # the Q table of ALL destinations is a projection of the layer input that every rank holds: made here
                qf_buf = (torch.empty if ld == d else torch.zeros)((h_full.shape[0], ld), dtype=dt, device=dev)
                _project(h_full.to(dt), w_q.to(dt), b_q, qf_buf, h_full.shape[0], d, ld)
                st["q_full"] = qf_buf
                del qf_buf
            else:
                if st["q_h"] is None:       # not prefetched by the layer above (cannot happen for l = L-1)
                    st["q_full"] = tr.new_full(q_sl, lease)
                    st["q_h"] = tr.gather(q_sl, st["q_full"])
                st["q_h"].wait()
            for hnd in da_handles:
                hnd.wait()
            _mark("bwd:wait_Q_dA")
            if l > 0 and not state[l - 1]["has_full"]:
                # the layer below needs its Q table next: it travels under this CSC walk
                sb = state[l - 1]
                sb["q_full"] = tr.new_full(sb["q_sl"], leases[l - 1])
                sb["q_h"] = tr.gather(sb["q_sl"], sb["q_full"])
            qf = _table(st["q_full"], st["q_full"].shape[0], d)
            sc_rows = part.scale_cols_rows(agg_type)
            w_cat = None
            if l > 0 or need[0]:
                w_cat = (torch.zeros if ld != d else torch.empty)((2 * ld, w_q.shape[1]), dtype=dt, device=dev)
                w_cat[:d].copy_(w_q)
                w_cat[ld:ld + d].copy_(w_k)
            g = None
            if l > 0 and ctx.bwd_chunks > 1:
                # dK walk in source chunks: chunk c of dH_l = [dQ|dK][c]·[W_Q;W_K] gives chunk c of the layer below's dA,
                # which travels while chunk c+1 is walked — the layer below then starts with its dA table in place
                sb, lb = state[l - 1], leases[l - 1]
                w_rb = weights[5 * (l - 1) + 3].to(dt)
                db_, ldb = sb["d"], sb["ld"]
                dsb = part.scales_rows(ctx.cfgs[l - 1][0])[0]
                dab_sl = lb.add(tr.acquire(part.n_pad, ldb, dt, dev, zero=ldb != db_))
                dab_full = tr.new_full(dab_sl, lb)
                g = torch.empty((n, w_q.shape[1]), dtype=dt, device=dev)
                handles = []
                for lo, hi, rows in part.col_chunks(ctx.bwd_chunks):
                    if rows is not None:
                        top = lo + rows.n_rows
                        be.backward_k(rows, qf, k[lo:top], None, daf, None, None if sc_rows is None else sc_rows[lo:top],
                                      act, act_param, out=dk[lo:top])
                        gemm.linear_dgrad(dqk[lo:top], w_cat, out=g[lo:top])
                        dab = dab_sl.local[lo:top, :db_]
                        if ldb == db_:
                            gemm.linear_dgrad(g[lo:top], w_rb, out=dab)
                        else:
                            dab.copy_(gemm.linear_dgrad(g[lo:top], w_rb))
                        if dsb is not None:
                            dab.mul_(dsb[lo:top].to(dt).unsqueeze(1))
                    if hi > lo:
                        handles.append(tr.gather(dab_sl, dab_full, lo, hi))
                prepared = (dab_sl, dab_full, handles)
                del dab_sl, dab_full
            else:
                be.backward_k(part.csc, qf, k, None, daf, None, sc_rows, act, act_param, out=dk)
                if w_cat is not None:
                    g = gemm.linear_dgrad(dqk, w_cat)
            _mark("bwd:edge_k")
            del da_full, daf, qf, kf, k_all, q, k, da, a, dq, dk, da_sl, q_sl, k_sl, da_handles, h_full
            for key in ("q_full", "q_h", "k_all", "a", "k_sl", "q_sl", "h_full", "q_all_kept"):
                st[key] = None
            featd = feat.to(dt)
            want_bq = bool(wneed[l][1] and b_q is not None)
            if wneed[l][0] or wneed[l][2]:
                dw_qk, db_qk = gemm.linear_wgrad_bias(dqk, featd, adt, want_bq)     # [2·ld, d_in], [2·ld]
                grads[l][0], grads[l][2] = dw_qk[:d], dw_qk[ld:ld + d]
                if db_qk is not None:
                    grads[l][1] = db_qk[:d]
            elif want_bq:
                grads[l][1] = gemm.column_sum(dqk, adt)[:d]
            st["h"] = None
            del feat, featd, dqk
            lease.release()                 # after the last reader of this layer's slices has been enqueued
            _mark("bwd:grads")
        # weight gradients of all layers: fp32 partials of every rank in ONE flat all-reduce
        live = [(l, i, t) for l in range(L) for i, t in enumerate(grads[l]) if t is not None and wneed[l][i]]
        out_w = [None] * (5 * L)
        if live:
            flat = torch.cat([t.reshape(-1) for _, _, t in live])
            part.all_reduce_(flat)
            off = 0
            for l, i, t in live:
                out_w[5 * l + i] = flat[off:off + t.numel()].view(t.shape).to(weights[5 * l + i].dtype)
                off += t.numel()
        _mark("bwd:allreduce")
        dfeat = g.to(feat0.dtype) if need[0] else None
        return (dfeat, None, None, None, None, None, *out_w)


def _layer_args(layer):
    from .conv import _SUM_LIKE, classify_activation
    known = classify_activation(layer.activation)
    if layer._agg_type not in _SUM_LIKE or known is None or not layer._plain():
        raise NotImplementedError("the partitioned path covers the fused configuration only "
                                  "(sum/mean/sym with ReLU/LeakyReLU/GELU/Identity)")
    if layer.training and layer.dropout.p > 0:
        raise NotImplementedError("dropout inside the partitioned layer is not supported")
    lq, lk, lr = layer.linear_query, layer.linear_key, layer.linear_relation
    return (layer._agg_type, known[0], known[1]), (lq.weight, lq.bias, lk.weight, lr.weight, lr.bias)


def partitioned_sirconv_stack(layers, part: RowPartition, feat_loc, chunks=4, backend=CudaEdgeBackend,
                              feat_full=None, gather="projections", bwd_chunks=1, keep_q_full=None):
    """Run consecutive `SIRConv` layers (output of one = input of the next, nothing in between) on this rank's rows
    of a partitioned graph as one autograd node; `chunks` = destination chunks of the cross-layer prefetch.
    `feat_loc` = rows [part.lo, part.hi) of the node features; `feat_full` (optional) = the rows of all ranks,
    part.all_gather_rows(feat_loc) — an input gathered ahead of time instead of two projections gathered in line.
    `gather`: what travels between layers — "projections" (the K and Q tables; default) or "inputs" (the layer
    input, projected locally by every rank: one table transfer per layer instead of two, paid for with two
    whole-table GEMMs per layer that nothing hides — worth it only where the link, not the dependency chain, is the
    limit).
    `bwd_chunks` > 1: the dK walk of every layer but the first runs in that many source chunks, and chunk c of the
    layer below's dA table (dH[c]·W_R, scaled) travels under the walk of chunk c+1 — the layer below then starts its
    dQ walk with no transfer competing for the SMs and its dA table already in place.
    `keep_q_full` (with feat_full / gather="inputs"): a layer whose gathered input is at hand makes its K AND Q tables
    of all ranks by ONE [Q|K] GEMM in forward and keeps Q for the backward CSC walk (None = when memory allows)."""
    if gather not in ("inputs", "projections"):
        raise ValueError(f"gather must be 'inputs' or 'projections', not {gather!r}")
    if feat_loc.shape[0] != part.n_local:
        raise ValueError(f"feat_loc has {feat_loc.shape[0]} rows, this rank owns {part.n_local}")
    if feat_full is not None and feat_full.shape[0] != part.world * part.n_pad:
        raise ValueError(f"feat_full has {feat_full.shape[0]} rows, expected {part.world * part.n_pad}")
    if feat_full is not None and feat_full.requires_grad:
        feat_full = feat_full.detach()
    cfgs, weights = [], []
    for layer in layers:
        c, w = _layer_args(layer)
        cfgs.append(c)
        weights += list(w)
    opts = {"chunks": chunks, "gather": gather, "bwd_chunks": bwd_chunks, "keep_q_full": keep_q_full}
    return PartitionedSIRStackFunction.apply(feat_loc, part, tuple(cfgs), opts, backend, feat_full, *weights)


def partitioned_sirconv(layer, part: RowPartition, feat_loc, backend=CudaEdgeBackend):
    """Run one `SIRConv` (sum / mean / sym, elementwise σ, no dropout) on this rank's rows of a partitioned graph."""
    return partitioned_sirconv_stack([layer], part, feat_loc, 1, backend, gather="projections")
