"""1-D destination-row partition of one large graph over the GPUs of a node (one process per GPU,
torch.distributed / NCCL over NVLink).  New capability — the reference is single-process — whose
correctness oracle is "G-rank result == 1-rank result" (SURVEY.md §8e).

Rank r owns node rows [r*n_pad, min(N, (r+1)*n_pad)), n_pad = ceil(N/G): its slice of H, Q, K, A, the
in-CSR rows of those destinations (column ids stay GLOBAL source ids) and the out-CSC rows of those
sources (row ids stay GLOBAL destination ids).  Per layer there are three exchanges, two of them hidden:

    forward   K_full  = all_gather(K_loc)                 -> edge_fwd over the local CSR rows   (exposed)
              Q_full  = all_gather(Q_loc)  issued right behind it on NCCL's stream; it is only needed by the
                        backward CSC pass, so it travels while the forward edge kernel runs
    backward  dA_loc *= dst coefficient;  dA_full = all_gather(dA_loc)  travels while
              dQ_loc  = edge_bwd_q over the local CSR rows (re-uses K_full) runs
              dK_loc  = edge_bwd_k over the local CSC rows (Q_full, dA_full, K_loc)
    weights   replicated; dW summed with all_reduce

Gathered tables are laid out [G*n_pad, ld], so a global node id indexes them directly.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist
import torch.nn.functional as F

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


def _phase_split(csr: CompressedRows, n_local, n_pad, world, phases, chunk):
    """(csr_k, [csr_phase_p]): the local in-CSR with source ids remapped to the [P, G, c] layout of the gathered K
    table, whole and split by source chunk (rows = local destinations in both)."""
    if phases == 1:
        return csr, [csr]
    idx = csr.idx.long()
    r = idx // n_pad
    i = idx - r * n_pad
    p = i // chunk
    remapped = ((p * world + r) * chunk + (i - p * chunk)).to(torch.int32)
    csr_k = CompressedRows(csr.indptr, remapped, None, csr.long_threshold)
    deg = (csr.indptr[1:] - csr.indptr[:-1]).long()
    dst = torch.repeat_interleave(torch.arange(n_local, dtype=torch.int64, device=idx.device), deg,
                                  output_size=int(idx.numel()))
    key = (p * n_local + dst).to(torch.int32)
    del r, i, dst, deg, idx
    if key.is_cuda:
        stacked = build_rows(key, remapped, phases * n_local, csr.long_threshold)
    else:       # host-side planning (gloo tests): same stable ordering with torch.sort
        order = torch.sort(key.long(), stable=True)[1]
        counts = torch.bincount(key.long(), minlength=phases * n_local)
        indptr = torch.zeros(phases * n_local + 1, dtype=torch.int32)
        indptr[1:] = torch.cumsum(counts, 0).to(torch.int32)
        stacked = CompressedRows(indptr, remapped[order].contiguous(), None, csr.long_threshold)
    return csr_k, [stacked.slice_rows(q * n_local, (q + 1) * n_local) for q in range(phases)]


class RowPartition:
    """Local slice of a graph for rank `rank` of `world` (see module docstring).

    csr: rows = local destinations, idx = GLOBAL sources;  csc: rows = local sources, idx = GLOBAL destinations.
    in_norm / out_norm / inv_in_deg: fp32 [G*n_pad] GLOBAL coefficient vectors (padding = 1)."""

    def __init__(self, num_nodes, rank, world, csr_local, csc_local, in_norm, out_norm, inv_in_deg, group=None,
                 phases=1):
        self.num_nodes_, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        self.n_pad = (self.num_nodes_ + world - 1) // world
        self.lo = min(self.num_nodes_, rank * self.n_pad)
        self.hi = min(self.num_nodes_, self.lo + self.n_pad)
        self.csr, self.csc = csr_local, csc_local
        self.in_norm, self.out_norm, self.inv_in_deg = in_norm, out_norm, inv_in_deg
        self.num_local_edges = csr_local.num_pos
        # Phased forward: every rank's K rows are cut into `phases` chunks of c rows; chunk p of ALL ranks is one
        # all-gather, and the forward edge pass runs once per chunk over the edges whose source lies in it, so the
        # walk over chunk p overlaps the transfer of chunks p+1..  The gathered table is laid out [P, G, c, ld]:
        # global id v = r*n_pad + p*c + j  ->  row (p*G + r)*c + j   (csr_k / csr_phase carry these remapped ids).
        self.phases = max(1, int(phases)) if world > 1 else 1
        self.chunk = (self.n_pad + self.phases - 1) // self.phases
        self.csr_k, self.csr_phase = _phase_split(self.csr, self.n_local, self.n_pad, world, self.phases, self.chunk)
        if self.phases == 1:
            self.out_norm_k = out_norm
        else:       # the per-source coefficient in the layout of the gathered K table
            P, c, G = self.phases, self.chunk, world
            v = out_norm.new_ones((G, P * c))
            v[:, :self.n_pad] = out_norm.view(G, self.n_pad)
            self.out_norm_k = v.view(G, P, c).transpose(0, 1).contiguous().view(-1)

    @property
    def k_rows(self):
        """rows of this rank's padded K buffer (= phases * chunk >= n_pad)"""
        return self.phases * self.chunk

    @staticmethod
    def bounds(num_nodes, rank, world):
        n_pad = (num_nodes + world - 1) // world
        lo = min(num_nodes, rank * n_pad)
        return n_pad, lo, min(num_nodes, lo + n_pad)

    @property
    def n_local(self):
        return self.hi - self.lo

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_csr_csc(cls, csr: CompressedRows, csc: CompressedRows, num_nodes, rank, world, group=None,
                     in_norm=None, out_norm=None, inv_in_deg=None, phases=1):
        """slice replicated whole-graph structures (tests, small graphs)"""
        n = int(num_nodes)
        n_pad, lo, hi = cls.bounds(n, rank, world)
        if in_norm is None:
            in_deg = (csr.indptr[1:] - csr.indptr[:-1]).clamp(min=1).to(torch.float32)
            out_deg = (csc.indptr[1:] - csc.indptr[:-1]).clamp(min=1).to(torch.float32)
            in_norm, out_norm, inv_in_deg = 1.0 / torch.sqrt(in_deg), 1.0 / torch.sqrt(out_deg), 1.0 / in_deg
        pad = lambda t: torch.cat([t, t.new_ones(world * n_pad - n)]) if world * n_pad > n else t
        return cls(n, rank, world, csr.slice_rows(lo, hi), csc.slice_rows(lo, hi),
                   pad(in_norm), pad(out_norm), pad(inv_in_deg), group, phases)

    @classmethod
    def from_graph(cls, graph: Graph, rank, world, group=None, phases=1):
        """slice an already converted (replicated) Graph; the caller may drop `graph` afterwards"""
        return cls.from_csr_csc(graph.csr, graph.csc, graph.num_nodes(), rank, world, group,
                                graph.in_norm, graph.out_norm, graph.inv_in_deg, phases)

    @classmethod
    def from_local_edges(cls, num_nodes, rank, world, in_src, in_dst, out_src, out_dst, group=None,
                         long_threshold=DEFAULT_LONG_THRESHOLD, in_indptr=None, phases=1):
        """build from this rank's two edge lists (GLOBAL ids): the edges whose destination is local
        (in_src -> in_dst) and the edges whose source is local (out_src -> out_dst).  When the first list is
        already destination-sorted, pass its local row pointer `in_indptr` instead of `in_dst`.  Degree
        coefficients of the whole graph are assembled with one all_gather each."""
        n = int(num_nodes)
        n_pad, lo, hi = cls.bounds(n, rank, world)
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
                   gather(1.0 / in_deg), group, phases)

    @classmethod
    def synthetic_powerlaw(cls, num_nodes, num_edges, rank, world, alpha=2.3, max_deg=None, seed=0, device="cuda",
                           group=None, long_threshold=DEFAULT_LONG_THRESHOLD, phases=1):
        """this rank's slice of synth.powerlaw_hashed (the SAME global graph on every world size), generated on
        the device without any edge exchange"""
        from . import synth
        indptr = synth.powerlaw_indptr(num_nodes, num_edges, alpha, max_deg, seed, device)
        n_pad, lo, hi = cls.bounds(num_nodes, rank, world)
        in_src, _ = synth.powerlaw_hashed_rows(indptr, num_nodes, lo, hi, seed, want_dst=False)
        out_src, out_dst = synth.powerlaw_hashed_cols(indptr, num_nodes, lo, hi, seed)
        in_indptr = indptr[lo:hi + 1] - indptr[lo]
        del indptr
        return cls.from_local_edges(num_nodes, rank, world, in_src, None, out_src, out_dst, group, long_threshold,
                                    in_indptr=in_indptr, phases=phases)

    # ---- coefficients --------------------------------------------------------------------------------------
    def _rows(self, v):
        return v[self.lo:self.lo + max(self.n_local, 1)]

    def scales_rows(self, agg_type):
        """(dst_scale, src_scale) for the CSR walks: dst = local row, src = row of the gathered K table"""
        if agg_type == "sym":
            return self._rows(self.in_norm), self.out_norm_k
        if agg_type == "mean":
            return self._rows(self.inv_in_deg), None
        return None, None

    def scale_cols_rows(self, agg_type):
        """row (= local source) scale of the CSC walk; its destination scale is folded into dA before the gather"""
        return self._rows(self.out_norm) if agg_type == "sym" else None

    # ---- collectives ------------------------------------------------------------------------------------------
    def new_rows(self, ld, dtype, device):
        """[n_pad, ld] buffer whose first n_local rows are this rank's slice of a table"""
        return torch.empty((self.n_pad, ld), dtype=dtype, device=device)

    def all_gather_rows(self, local_pad, async_op=False):
        """[n_pad, ld] -> ([G*n_pad, ld], work handle or None); rows of rank r land at r*n_pad"""
        full = local_pad.new_empty((self.world * self.n_pad, local_pad.shape[1]))
        if self.world == 1:
            full.copy_(local_pad)
            return full, None
        work = dist.all_gather_into_tensor(full, local_pad, group=self.group, async_op=async_op)
        return full, (work if async_op else None)

    def all_gather_k(self, k_pad):
        """phased gather of K: k_pad [P*c, ld] -> (K_all [P*G*c, ld], [work per chunk]); chunk p of every rank
        lands in K_all[p*G*c : (p+1)*G*c]"""
        P, c, G = self.phases, self.chunk, self.world
        k_all = k_pad.new_empty((P * G * c, k_pad.shape[1]))
        if G == 1:
            k_all.copy_(k_pad)
            return k_all, [None]
        works = [dist.all_gather_into_tensor(k_all[p * G * c:(p + 1) * G * c], k_pad[p * c:(p + 1) * c],
                                             group=self.group, async_op=True) for p in range(P)]
        return k_all, works

    def all_reduce_(self, t):
        if self.world > 1 and t is not None:
            dist.all_reduce(t, group=self.group)
        return t


# bench.py sets this to a list to collect (label, CUDA event) marks on the compute stream (phase breakdown)
PHASE_MARKS = None


def _mark(label):
    if PHASE_MARKS is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        PHASE_MARKS.append((label, ev))


def _wait(work):
    if work is not None:
        work.wait()        # stream-level wait: the current stream waits for NCCL's, the host does not block


class PartitionedSIRLayerFunction(torch.autograd.Function):
    """SIRLayerFunction for a row-partitioned graph: same arithmetic per row; K all-gathered in forward, Q and
    the pre-scaled dA all-gathered behind the edge kernels, weight gradients all-reduced (every rank ends with
    the full-graph gradient)."""

    @staticmethod
    def forward(ctx, feat, w_q, b_q, w_k, w_r, b_r, part: RowPartition, agg_type, act, act_param, backend):
        n, d, dt, dev = part.n_local, w_q.shape[0], feat.dtype, feat.device
        ld = F_._pad_cols(d, dt)
        alloc = (torch.zeros if ld != d else torch.empty)
        _mark("fwd:start")
        k_pad = alloc((part.k_rows, ld), dtype=dt, device=dev)
        q_pad = alloc((part.n_pad, ld), dtype=dt, device=dev)
        wq, wk = w_q.to(dt), w_k.to(dt)
        if ld == d:
            gemm.linear_forward(feat, wk, None, out=k_pad[:n])
        else:
            k_pad[:n, :d].copy_(gemm.linear_forward(feat, wk, None))
        k_all, k_works = part.all_gather_k(k_pad)                       # chunk by chunk, in phase order
        if ld == d:
            gemm.linear_forward(feat, wq, b_q, out=q_pad[:n])
        else:
            q_pad[:n, :d].copy_(gemm.linear_forward(feat, wq, b_q))
        q_full, wq_h = part.all_gather_rows(q_pad, async_op=True)       # consumed by backward only
        q, k = q_pad[:n, :d], k_pad[:n, :d]
        q._sirgcn_padded = k._sirgcn_padded = True
        kf = k_all[:, :d]
        kf._sirgcn_padded = True
        ds, ss = part.scales_rows(agg_type)
        _mark("fwd:proj")
        a = F_._alloc_table(n, d, dt, dev, zero=ld != d)
        for p, work in enumerate(k_works):                              # walk chunk p while chunks p+1.. travel
            _wait(work)
            if p == 0:
                _mark("fwd:wait_K0")
            backend.forward(part.csr_phase[p], q, kf, None, ds, ss, act, act_param, out=a, accumulate=p > 0)
        a._sirgcn_padded = True
        _mark("fwd:edge")
        out = gemm.linear_forward(a, w_r.to(dt), b_r)
        _mark("fwd:out")
        ctx.save_for_backward(feat, q_pad, k_pad, k_all, q_full, a, w_q, w_k, w_r)
        ctx.q_work = wq_h
        ctx.part, ctx.agg_type, ctx.act, ctx.act_param, ctx.backend = part, agg_type, act, act_param, backend
        ctx.has_bias = (b_q is not None, b_r is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        feat, q_pad, k_pad, k_all, q_full, a, w_q, w_k, w_r = ctx.saved_tensors
        part, be = ctx.part, ctx.backend
        n, d, dt, dev = part.n_local, w_q.shape[0], q_pad.dtype, q_pad.device
        ld = q_pad.shape[1]
        need = ctx.needs_input_grad
        gout = gout.to(dt)
        gout = gout if gout.stride(-1) == 1 else gout.contiguous()
        _mark("bwd:start")
        adt = torch.float64 if dt == torch.float64 else torch.float32      # dtype of the reduced weight gradients
        dw_r = (gout.t() @ a).to(adt) if need[4] else None
        db_r = gemm.column_sum(gout, adt) if (need[5] and ctx.has_bias[1]) else None
        # dA, scaled by the destination coefficient BEFORE it travels: the CSC pass then needs no scale lookup
        alloc = (torch.zeros if ld != d else torch.empty)
        da_pad = alloc((part.n_pad, ld), dtype=dt, device=dev)
        da = da_pad[:n, :d]
        if ld == d:
            gemm.linear_dgrad(gout, w_r.to(dt), out=da)
        else:
            da.copy_(gemm.linear_dgrad(gout, w_r.to(dt)))
        ds, ss = part.scales_rows(ctx.agg_type)
        if ds is not None:
            da.mul_(ds[:n].to(dt).unsqueeze(1))
        da._sirgcn_padded = True
        da_full, wa_h = part.all_gather_rows(da_pad, async_op=True)      # travels while dQ is computed
        q, k = q_pad[:n, :d], k_pad[:n, :d]
        q._sirgcn_padded = k._sirgcn_padded = True
        kf, qf, daf = k_all[:, :d], q_full[:, :d], da_full[:, :d]
        kf._sirgcn_padded = qf._sirgcn_padded = daf._sirgcn_padded = True
        dq = F_._alloc_table(n, d, dt, dev, zero=ld != d)
        dk = F_._alloc_table(n, d, dt, dev, zero=ld != d)
        _mark("bwd:dA")
        be.backward_q(part.csr_k, q, kf, None, da, None, ss, ctx.act, ctx.act_param, False, out=dq)
        _mark("bwd:edge_q")
        _wait(ctx.q_work)
        _wait(wa_h)
        _mark("bwd:wait_Q_dA")
        be.backward_k(part.csc, qf, k, None, daf, None, part.scale_cols_rows(ctx.agg_type), ctx.act, ctx.act_param,
                      out=dk)
        _mark("bwd:edge_k")
        featd = feat.to(dt)
        # weight gradients: fp32 partials of every rank in ONE flat all-reduce
        parts = [(dq.t() @ featd).to(adt) if need[1] else None,
                 gemm.column_sum(dq, adt) if (need[2] and ctx.has_bias[0]) else None,
                 (dk.t() @ featd).to(adt) if need[3] else None, dw_r, db_r]
        live = [t for t in parts if t is not None]
        if live:
            flat = torch.cat([t.reshape(-1) for t in live])
            part.all_reduce_(flat)
            off = 0
            for t in live:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
        dw_q, db_q, dw_k, dw_r, db_r = [None if t is None else t.to(w.dtype)
                                        for t, w in zip(parts, (w_q, w_q, w_k, w_r, w_r))]
        dfeat = None
        if need[0]:
            dfeat = gemm.linear_dgrad(dq, w_q.to(dt))
            dfeat.add_(gemm.linear_dgrad(dk, w_k.to(dt)))
            dfeat = dfeat.to(feat.dtype)
        _mark("bwd:grads")
        return dfeat, dw_q, db_q, dw_k, dw_r, db_r, None, None, None, None, None


def partitioned_sirconv(layer, part: RowPartition, feat_loc, backend=CudaEdgeBackend):
    """Run a `SIRConv` (sum / mean / sym, elementwise σ, no dropout) on this rank's rows of a
    partitioned graph.  `feat_loc` = rows [part.lo, part.hi) of the node features."""
    from .conv import _SUM_LIKE, classify_activation
    known = classify_activation(layer.activation)
    if layer._agg_type not in _SUM_LIKE or known is None or not layer._plain():
        raise NotImplementedError("the partitioned path covers the fused configuration only "
                                  "(sum/mean/sym with ReLU/LeakyReLU/GELU/Identity)")
    if layer.training and layer.dropout.p > 0:
        raise NotImplementedError("dropout inside the partitioned layer is not supported")
    if feat_loc.shape[0] != part.n_local:
        raise ValueError(f"feat_loc has {feat_loc.shape[0]} rows, this rank owns {part.n_local}")
    lq, lk, lr = layer.linear_query, layer.linear_key, layer.linear_relation
    return PartitionedSIRLayerFunction.apply(feat_loc, lq.weight, lq.bias, lk.weight, lr.weight, lr.bias, part,
                                             layer._agg_type, known[0], known[1], backend)
