"""Graph index structures consumed by the CUDA edge kernels.

The reference hands a ``DGLGraph`` to the layer (/root/reference/models/conv.py:49) and lets DGL
materialise the in-CSR for ``update_all`` and the degree vectors (conv.py:51-52, :63).  Here a graph
is converted ONCE into a destination-sorted CSR and a source-sorted CSC of int32 indices (stable:
ties keep edge-id order) by ``sirgcn_csr_build`` on the GPU, together with the degree norms and the
long-row schedule; the result is cached on the object.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib

DEFAULT_LONG_THRESHOLD = 256


class CompressedRows:
    """One walkable compressed-row structure (the in-CSR or the out-CSC, or a row slice of one)
    plus its long-row schedule.  ``idx`` holds the other endpoint of each stored edge, ``eid`` the
    original edge id (or None)."""

    def __init__(self, indptr, idx, eid, long_threshold=DEFAULT_LONG_THRESHOLD, sched=None, counts=None):
        self.indptr, self.idx, self.eid = indptr, idx, eid
        self.n_rows = int(indptr.numel()) - 1
        self.num_pos = int(idx.numel())
        self.long_threshold = int(long_threshold)
        if sched is None:
            if indptr.is_cuda:
                sched, counts = _build_schedule(indptr, self.n_rows, self.num_pos, self.long_threshold)
            else:   # host-side structure (partition planning / gloo tests): never handed to a kernel
                sched, counts = [torch.empty(0, dtype=torch.int32)] * 6, (0, 0)
        self._sched_tensors = sched
        self.n_long, self.n_chunks = int(counts[0]), int(counts[1])
        self.sched = _lib.Schedule(*[_lib.ptr(t).value for t in sched])
        self._partial = {}
        self.tile_row, self.n_tiles = _build_tiles(indptr, self.n_rows, self.num_pos)

    def partial(self, d, dtype):
        """fp32 scratch for the long-row partial sums (cached per (d, dtype))."""
        if self.n_chunks == 0:
            return None
        key = (d, dtype)
        buf = self._partial.get(key)
        if buf is None:
            nbytes = _lib.lib().sirgcn_edge_partial_bytes(self.n_chunks, d, _lib.DTYPE_CODE[dtype])
            buf = torch.empty(nbytes // 4, dtype=torch.float32, device=self.indptr.device)
            self._partial = {key: buf}
        return buf

    def slice_rows(self, lo, hi):
        """Rows [lo, hi) as a stand-alone structure (idx keeps global ids) — used by the
        destination-row partition (partition.py)."""
        beg, end = int(self.indptr[lo]), int(self.indptr[hi])
        indptr = (self.indptr[lo:hi + 1] - beg).contiguous()
        eid = None if self.eid is None else self.eid[beg:end].contiguous()
        return CompressedRows(indptr, self.idx[beg:end].contiguous(), eid, self.long_threshold)


def _build_tiles(indptr, n_rows, num_pos):
    """work tiles of the edge kernels (sirgcn.h "Work tiles"): (tile_row int32 [n_tiles+1], n_tiles)"""
    if not indptr.is_cuda or n_rows == 0:
        return None, 0
    L = _lib.lib()
    n_tiles = int(L.sirgcn_num_tiles(C.c_int32(n_rows), C.c_int64(num_pos)))
    tile_row = torch.empty(n_tiles + 1, dtype=torch.int32, device=indptr.device)
    with torch.cuda.device(indptr.device):
        _lib.check(L.sirgcn_tiles_build(_lib.ptr(indptr), C.c_int32(n_rows), C.c_int64(num_pos),
                                        _lib.ptr(tile_row), _lib.stream_ptr(indptr.device)), "sirgcn_tiles_build")
    return tile_row, n_tiles


def _sched_alloc(num_pos, thr, device):
    cap_long = num_pos // thr + 1
    cap_chunks = 2 * (num_pos // thr) + 2
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device=device)
    cap_big = num_pos // (thr * 32) + 2            # SIRGCN_BIG_CHUNKS
    return [i32(cap_long), i32(cap_long), i32(cap_long), i32(cap_chunks), i32(cap_chunks), i32(cap_big + 1)]


def _build_schedule(indptr, n_rows, num_pos, thr):
    dev = indptr.device
    sched = _sched_alloc(num_pos, thr, dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    s = _lib.Schedule(*[_lib.ptr(t).value for t in sched])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sirgcn_schedule_build(_lib.ptr(indptr), C.c_int32(n_rows), C.c_int32(thr),
                                                    C.byref(s), _lib.ptr(counts), _lib.stream_ptr(dev)),
                   "sirgcn_schedule_build")
    return sched, counts.tolist()


class Graph:
    """COO edge list src->dst over ``num_nodes`` nodes, converted on construction.

    Attributes: ``csr`` (rows = destinations, idx = sources), ``csc`` (rows = sources, idx =
    destinations), ``in_norm``/``out_norm`` = clamp(deg,1)^-1/2, ``inv_in_deg`` = 1/clamp(in_deg,1).
    """

    def __init__(self, src, dst, num_nodes, *, need_eid=True, long_threshold=DEFAULT_LONG_THRESHOLD,
                 validate=False, keep_coo=True):
        if not (torch.is_tensor(src) and src.is_cuda):
            raise RuntimeError("Graph needs CUDA index tensors: the SIR-GCN kernels have no CPU fallback")
        dev = src.device
        src = src.to(torch.int32).contiguous()
        dst = dst.to(device=dev, dtype=torch.int32).contiguous()
        E, N = int(src.numel()), int(num_nodes)
        if dst.numel() != E:
            raise ValueError("src and dst differ in length")
        if E >= 2 ** 31:
            raise ValueError("num_edges must be < 2^31 (int32 indices)")
        if validate and E > 0:
            lo = min(int(src.min()), int(dst.min()))
            hi = max(int(src.max()), int(dst.max()))
            if lo < 0 or hi >= N:
                raise ValueError(f"node ids out of range [0, {N}): min={lo} max={hi}")
        self.num_nodes_, self.num_edges_, self.device = N, E, dev
        thr = int(long_threshold)
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
        f32 = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
        indptr_in, col_src, indptr_out, row_dst = i32(N + 1), i32(E), i32(N + 1), i32(E)
        eid_in = i32(E) if need_eid else None
        eid_out = i32(E) if need_eid else None
        self.in_norm, self.out_norm, self.inv_in_deg = f32(N), f32(N), f32(N)
        sched_in, sched_out = _sched_alloc(E, thr, dev), _sched_alloc(E, thr, dev)
        s_in = _lib.Schedule(*[_lib.ptr(t).value for t in sched_in])
        s_out = _lib.Schedule(*[_lib.ptr(t).value for t in sched_out])
        counts = torch.zeros(5, dtype=torch.int32, device=dev)
        L = _lib.lib()
        ws_bytes = L.sirgcn_csr_build_workspace_bytes(E, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.sirgcn_csr_build(
                _lib.ptr(src), _lib.ptr(dst), C.c_int64(E), C.c_int32(N),
                _lib.ptr(indptr_in), _lib.ptr(col_src), _lib.ptr(eid_in),
                _lib.ptr(indptr_out), _lib.ptr(row_dst), _lib.ptr(eid_out),
                _lib.ptr(self.in_norm), _lib.ptr(self.out_norm), _lib.ptr(self.inv_in_deg),
                C.c_int32(thr), C.byref(s_in), C.byref(s_out), _lib.ptr(counts),
                _lib.ptr(ws), C.c_size_t(ws_bytes), _lib.stream_ptr(dev))
        _lib.check(rc, "sirgcn_csr_build")
        c = counts.tolist()  # the one host sync per graph: sizes of the long-row schedule (+ the id range flag)
        del ws
        if c[4]:
            raise ValueError(f"node ids out of range [0, {N}) in src/dst")
        self.csr = CompressedRows(indptr_in, col_src, eid_in, thr, sched_in, c[0:2])
        self.csc = CompressedRows(indptr_out, row_dst, eid_out, thr, sched_out, c[2:4])
        self.src, self.dst = (src, dst) if keep_coo else (None, None)
        self._lazy = {}

    # --- edge-subset views: the DropEdge fast path (models/utils.py:96-102) ---------------------
    def edge_subgraph(self, keep):
        """(sub, new_id): the graph restricted to the edges with keep[e] != 0, converted WITHOUT sorting
        (sirgcn_edge_subgraph: compaction of this graph's stable CSR / CSC).  Kept edges are renumbered in edge-id
        order — exactly what DGL's remove_edges does — so per-edge features of the sub-graph are `efeat[keep]`;
        new_id[e] is the new id of kept edge e.  Bit-identical to Graph(src[keep], dst[keep], N)."""
        if self.csr.eid is None:
            raise RuntimeError("edge_subgraph needs a graph built with need_eid=True")
        E, N, dev = self.num_edges_, self.num_nodes_, self.device
        keep = keep.to(device=dev)
        if keep.numel() != E:
            raise ValueError(f"keep has {keep.numel()} entries, the graph has {E} edges")
        keep8 = (keep != 0).to(torch.uint8).contiguous()
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
        f32 = lambda n: torch.empty(n, dtype=torch.float32, device=dev)
        new_id = i32(E + 1)
        ip_in, idx_in, eid_in, ip_out, idx_out, eid_out = i32(N + 1), i32(E), i32(E), i32(N + 1), i32(E), i32(E)
        sub = object.__new__(Graph)
        sub.in_norm, sub.out_norm, sub.inv_in_deg = f32(N), f32(N), f32(N)
        L = _lib.lib()
        nbytes = L.sirgcn_edge_subgraph_workspace_bytes(C.c_int64(E))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        c, o = self.csr, self.csc
        with torch.cuda.device(dev):
            rc = L.sirgcn_edge_subgraph(
                _lib.ptr(keep8), C.c_int64(E), C.c_int32(N),
                _lib.ptr(c.indptr), _lib.ptr(c.idx), _lib.ptr(c.eid), _lib.ptr(o.indptr), _lib.ptr(o.idx), _lib.ptr(o.eid),
                _lib.ptr(new_id), _lib.ptr(ip_in), _lib.ptr(idx_in), _lib.ptr(eid_in),
                _lib.ptr(ip_out), _lib.ptr(idx_out), _lib.ptr(eid_out),
                _lib.ptr(sub.in_norm), _lib.ptr(sub.out_norm), _lib.ptr(sub.inv_in_deg),
                _lib.ptr(ws), C.c_size_t(nbytes), _lib.stream_ptr(dev))
        _lib.check(rc, "sirgcn_edge_subgraph")
        kept = int(new_id[E])               # the one host sync: how many edges stay
        del ws
        thr = c.long_threshold
        sub.num_nodes_, sub.num_edges_, sub.device = N, kept, dev
        sub.csr = CompressedRows(ip_in, idx_in[:kept], eid_in[:kept], thr)
        sub.csc = CompressedRows(ip_out, idx_out[:kept], eid_out[:kept], thr)
        if self.src is not None:
            sel = keep8.bool()
            sub.src, sub.dst = self.src[sel], self.dst[sel]
        else:
            sub.src = sub.dst = None
        sub._lazy = {}
        return sub, new_id[:E]

    def drop_edges(self, p, generator=None):
        """DropEdge(p) on the converted graph (dgl.transforms.DropEdge semantics: every edge is removed
        independently with probability p): returns (sub-graph, keep mask) — index per-edge features with the mask.
        p = 0 returns this graph itself: nothing to convert."""
        E, dev = self.num_edges_, self.device
        if p <= 0 or E == 0:
            return self, torch.ones(E, dtype=torch.bool, device=dev)
        keep = torch.rand(E, device=dev, generator=generator) >= p
        sub, _ = self.edge_subgraph(keep)
        return sub, keep

    # --- the slice of the DGLGraph API the layer relies on (conv.py:50-55) ---------------------
    def num_nodes(self):
        return self.num_nodes_

    def num_edges(self):
        return self.num_edges_

    def in_degrees(self):
        return (self.csr.indptr[1:] - self.csr.indptr[:-1]).long()

    def out_degrees(self):
        return (self.csc.indptr[1:] - self.csc.indptr[:-1]).long()

    # --- helpers of the split (generic σ / max) path ------------------------------------------
    def pos_dst(self):
        """destination node of every CSR position, int32 [E]"""
        t = self._lazy.get("pos_dst")
        if t is None:
            deg = self.csr.indptr[1:] - self.csr.indptr[:-1]
            t = torch.repeat_interleave(torch.arange(self.num_nodes_, dtype=torch.int32, device=self.device),
                                        deg.long(), output_size=self.num_edges_)
            self._lazy["pos_dst"] = t
        return t

    def pos_eid(self):
        """edge id of every CSR position, int32 [E]"""
        if self.csr.eid is None:
            raise RuntimeError("graph was built with need_eid=False")
        return self.csr.eid

    def csc_pos(self):
        """CSR position of the edge stored at every CSC position (int32 [E])"""
        if self.csr.eid is None:
            raise RuntimeError("graph was built with need_eid=False")
        t = self._lazy.get("csc_pos")
        if t is None:
            inv = torch.empty(self.num_edges_, dtype=torch.int32, device=self.device)
            inv[self.csr.eid.long()] = torch.arange(self.num_edges_, dtype=torch.int32, device=self.device)
            t = inv[self.csc.eid.long()].contiguous()
            self._lazy["csc_pos"] = t
        return t

    def edge_type_positions(self, etypes):
        """(type of the edge stored at every CSR position, ... at every CSC position), int32 [E] each, for an integer
        per-edge-id tensor `etypes` — what the kernels index a small edge-term TABLE with (12 bytes of traffic per
        edge instead of an [E, d] projected edge tensor).  Cached for the tensor last seen (all layers of a model use
        the same bond types)."""
        if self.csr.eid is None:
            raise RuntimeError("edge features need a graph built with need_eid=True")
        if etypes.dim() != 1 or etypes.shape[0] != self.num_edges_ or etypes.is_floating_point():
            raise ValueError(f"edge types must be an integer tensor of shape [{self.num_edges_}]")
        key = (etypes.data_ptr(), etypes._version, etypes.dtype)
        hit = self._lazy.get("etype_pos")
        if hit is not None and hit[0] == key and hit[1] is etypes:
            return hit[2], hit[3]
        t32 = etypes.to(device=self.device, dtype=torch.int32)
        ix_csr, ix_csc = t32[self.csr.eid.long()].contiguous(), t32[self.csc.eid.long()].contiguous()
        self._lazy["etype_pos"] = (key, etypes, ix_csr, ix_csc)
        return ix_csr, ix_csc

    def scales(self, agg_type):
        """(dst_scale, src_scale) fp32 vectors realising the aggregator coefficient c_e of
        conv.py:45 / fn.mean: sum -> (None, None); mean -> (1/clamp(in_deg,1), None);
        sym -> (in_deg^-1/2, out_deg^-1/2) with degrees clamped to >= 1 (conv.py:51-57)."""
        if agg_type == "sym":
            return self.in_norm, self.out_norm
        if agg_type == "mean":
            return self.inv_in_deg, None
        return None, None


class DropEdge:
    """Call-compatible stand-in for the reference's `DropEdge(p)(graph, efeats) -> (graph, efeats)`
    (/root/reference/models/utils.py:96-102, a dgl.transforms.DropEdge that also subsets the edge features), on this
    package's converted graphs: no rebuild, no sort (Graph.drop_edges), and p = 0 hands the graph back untouched."""

    def __init__(self, p=0.5):
        self.p = float(p)

    def __call__(self, graph, efeats=None):
        g = as_graph(graph)
        sub, keep = g.drop_edges(self.p)
        if efeats is None or sub is g:
            return sub, efeats
        return sub, efeats[keep]


_dgl_cache = weakref.WeakKeyDictionary()


def as_graph(graph) -> Graph:
    """Accept the repo's own Graph or (when DGL is installed) a DGLGraph, converted once and
    cached by object identity; the caller's graph is never mutated (conv.py:50 local_scope)."""
    if isinstance(graph, Graph):
        return graph
    if hasattr(graph, "edges") and hasattr(graph, "num_nodes"):
        try:
            cached = _dgl_cache.get(graph)
        except TypeError:
            cached = None
        if cached is not None:
            return cached
        src, dst = graph.edges(form="uv", order="eid")     # int64 by default in DGL; Graph narrows to int32
        if not src.is_cuda:
            raise RuntimeError("the DGLGraph lives on the host: move it to the GPU first (`graph.to(device)`, as the "
                               "reference scripts do) — the SIR-GCN kernels need CUDA tensors and have no CPU fallback")
        g = Graph(src, dst, graph.num_nodes())
        try:
            _dgl_cache[graph] = g
        except TypeError:
            pass
        return g
    raise TypeError(f"unsupported graph type {type(graph)!r}")
