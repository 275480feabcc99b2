"""torch.autograd.Function wrappers around the C-ABI edge kernels.

``EdgeAggregate`` is the fused replacement of ``graph.update_all(message_func, fn.sum|mean)`` and
its autograd backward (/root/reference/models/conv.py:43-47,:63): it saves Q, K (and the caller's
projected edge term) — never an |E| x d tensor — and recomputes the pre-activation in backward.
``GatherAdd`` / ``SegmentReduce`` form the split path for arbitrary σ, max/min and the *Base layers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .graph import CompressedRows, Graph


def _pad_cols(d, dtype):
    """columns of a 16-byte aligned row holding d elements"""
    per = 16 // torch.empty((), dtype=dtype).element_size()
    return (d + per - 1) // per * per


def _alloc_table(n, d, dtype, device, zero=False):
    """[n, d] view of a row-padded buffer (rows 16-B aligned)."""
    ld = _pad_cols(d, dtype)
    buf = (torch.zeros if zero else torch.empty)((n, ld), dtype=dtype, device=device)
    return buf[:, :d]


def copy_rows(dst, src):
    """dst[:, :] = src for two row tables of the same shape and dtype whose rows are 16-byte aligned (any row
    strides): sirgcn_copy_rows when the row is a whole number of 16-byte vectors, torch otherwise"""
    es = src.element_size()
    rb = src.shape[1] * es
    ok = (src.is_cuda and dst.is_cuda and src.dtype == dst.dtype and src.shape == dst.shape and src.dim() == 2
          and src.stride(1) == 1 and dst.stride(1) == 1 and rb % 16 == 0 and src.data_ptr() % 16 == 0
          and dst.data_ptr() % 16 == 0 and (src.stride(0) * es) % 16 == 0 and (dst.stride(0) * es) % 16 == 0)
    if not ok or src.shape[0] == 0:
        return dst.copy_(src)
    with torch.cuda.device(src.device):
        rc = _lib.lib().sirgcn_copy_rows(_lib.ptr(dst), C.c_int64(dst.stride(0) * es), _lib.ptr(src),
                                         C.c_int64(src.stride(0) * es), C.c_int64(rb), C.c_int64(src.shape[0]),
                                         _lib.stream_ptr(src.device))
    _lib.check(rc, "sirgcn_copy_rows")
    return dst


def mask_scale_(table, keep, scale):
    """in-place dropout on a row table view [n, d] (16-byte aligned, padded rows): table = keep ? table·scale : 0
    (sirgcn_mask_scale; keep = torch.bool [n, d] contiguous)"""
    if table.shape[0] == 0:
        return table
    assert keep.dtype == torch.bool and keep.is_contiguous() and keep.shape == table.shape
    es = table.element_size()
    with torch.cuda.device(table.device):
        rc = _lib.lib().sirgcn_mask_scale(_lib.ptr(table), C.c_int64(_ld(table) * es), _lib.ptr(keep),
                                          C.c_int64(table.shape[0]), C.c_int32(table.shape[1]),
                                          C.c_int32(_lib.DTYPE_CODE[table.dtype]), C.c_float(scale),
                                          _lib.stream_ptr(table.device))
    _lib.check(rc, "sirgcn_mask_scale")
    return table


def draw_keep_mask(n, d, dtype, device, p):
    """the keep mask nn.Dropout(p) would draw for a fresh contiguous [n, d] tensor of `dtype` in training mode, with
    the framework's own generator (same Philox consumption: ATen's fused dropout keys its random stream on numel,
    dtype and alignment only) — so a seeded run draws the very masks the reference's `self.dropout(...)` calls draw,
    in whatever order the caller asks for them (/root/reference/models/conv.py:60-61,:128: K, Q, E)"""
    if p >= 1:
        return torch.zeros((n, d), dtype=torch.bool, device=device)      # nn.Dropout(1): zeros, no random numbers
    _, keep = torch.native_dropout(torch.empty((n, d), dtype=dtype, device=device), p, True)
    return keep


def _table_ok(t):
    es = t.element_size()
    if t.dim() != 2 or t.stride(1) != 1 or t.data_ptr() % 16:
        return False
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), _pad_cols(t.shape[1], t.dtype))
    if (ld * es) % 16 or ld < _pad_cols(t.shape[1], t.dtype):
        return False
    # columns d..pad must be readable zeros: only guaranteed for buffers made by _alloc_table /
    # the padded GEMM, which is the case when the row is already a whole number of 16-B vectors
    return (t.shape[1] * es) % 16 == 0 or getattr(t, "_sirgcn_padded", False)


def as_table(t):
    """Return a tensor the kernels can read as a table (16-B aligned rows, zero padding)."""
    if _table_ok(t):
        return t
    out = _alloc_table(t.shape[0], t.shape[1], t.dtype, t.device, zero=True)
    out.copy_(t)
    out._sirgcn_padded = True
    return out


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else _pad_cols(t.shape[1], t.dtype)


# Persistent walks (a grid of the resident CTAs + a device work-unit counter) are the default; the row-partitioned stack
# (partition.py) switches to plain grids while it runs at more than one rank, so that the collectives' kernels can become
# resident under a walk (include/sirgcn.h SIRGCN_WALK_PLAIN_GRID).
WALK_PERSISTENT = True

# bench.py sets this to a list to collect (entry point, start event, end event, (positions, rows)) of every edge
# call, recorded on the stream the kernels are launched on (roofline measurement); None = off
EDGE_TIMERS = None


def _edge_call(fn_name, rows: CompressedRows, d, dtype, act, act_param, q, k, da, e, out, de,
               dst_scale, src_scale, da_scaled=None, accumulate=False, e_index=None, de_partial=None):
    """`e_index` (int32 per stored position): the edge term is row e_index[p] of the small TABLE `e` instead of row
    eid[p] of an [E, d] tensor; `de_partial` (zeroed fp32 scratch): the dQ walk reduces that table's gradient."""
    if not (rows.indptr.is_cuda and q.is_cuda and k.is_cuda and out.is_cuda):
        raise RuntimeError("SIR-GCN edge kernels need CUDA tensors (no CPU fallback)")
    a = _lib.EdgeArgs()
    a.n_rows, a.d, a.dtype, a.act = rows.n_rows, d, _lib.DTYPE_CODE[dtype], act
    a.act_param, a.long_threshold = float(act_param), rows.long_threshold
    # an edgeless graph has an empty idx tensor (NULL data_ptr): hand the kernels any valid address,
    # it is never dereferenced because every row is empty
    a.indptr, a.idx = rows.indptr.data_ptr(), (rows.idx.data_ptr() or rows.indptr.data_ptr())
    a.eid = e_index.data_ptr() if e_index is not None else (None if rows.eid is None else rows.eid.data_ptr())
    a.q, a.ldq = q.data_ptr(), _ld(q)
    a.k, a.ldk = k.data_ptr(), _ld(k)
    if da is not None:
        a.da, a.lda = da.data_ptr(), _ld(da)
    if e is not None:
        a.e, a.lde = e.data_ptr(), _ld(e)
    a.out, a.ldo = out.data_ptr(), _ld(out)
    if de is not None:
        a.de, a.ldde = de.data_ptr(), _ld(de)
    if da_scaled is not None:
        a.da_scaled, a.ldds = da_scaled.data_ptr(), _ld(da_scaled)
    a.dst_scale = None if dst_scale is None else dst_scale.data_ptr()
    a.src_scale = None if src_scale is None else src_scale.data_ptr()
    a.sched, a.n_long, a.n_chunks = rows.sched, rows.n_long, rows.n_chunks
    partial = rows.partial(d, dtype)
    a.partial = None if partial is None else partial.data_ptr()
    a.tile_row = None if rows.tile_row is None else rows.tile_row.data_ptr()
    a.n_tiles = rows.n_tiles
    a.accumulate = 1 if accumulate else 0
    a.flags = 0 if WALK_PERSISTENT else _lib.WALK_PLAIN_GRID
    if de_partial is not None:
        a.n_etypes, a.de_partial = e.shape[0], de_partial.data_ptr()
    dev = rows.indptr.device
    with torch.cuda.device(dev):
        if EDGE_TIMERS is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
        rc = getattr(_lib.lib(), fn_name)(C.byref(a), _lib.stream_ptr(dev))
        if EDGE_TIMERS is not None:
            t1.record()
            EDGE_TIMERS.append((fn_name, t0, t1, (rows.num_pos, rows.n_rows)))   # sizes only: never keep a graph alive
    _lib.check(rc, fn_name)


def edge_forward(csr: CompressedRows, q, k, e, dst_scale, src_scale, act, act_param, out=None, accumulate=False,
                 e_index=None):
    """A = fused edge stage over the rows of `csr` (q indexed by row, k by csr.idx); with `accumulate`,
    `out` (+)= the result — one walk per source block of a phased, partitioned forward."""
    d = q.shape[1]
    if out is None:
        out = _alloc_table(csr.n_rows, d, q.dtype, q.device)
    if csr.n_rows:
        _edge_call("sirgcn_edge_fwd", csr, d, q.dtype, act, act_param, q, k, None, e, out, None,
                   dst_scale, src_scale, accumulate=accumulate, e_index=e_index)
    out._sirgcn_padded = True
    return out


def edge_backward_q(csr, q, k, e, da, dst_scale, src_scale, act, act_param, want_de, out=None,
                    scale_da_inplace=False, e_index=None):
    """dQ over the CSR (+ dE).  With `scale_da_inplace` (and a destination scale), the pass also overwrites
    dA[u] by dst_scale[u]·dA[u], so the CSC pass that follows can gather an already scaled table and needs
    no per-edge scale lookup (pass it dst_scale=None).  With `e_index` (edge term = rows of the small table `e`)
    the returned dE is the TABLE's gradient, fp32 [n_types, d], reduced inside the walk: no [E, d] tensor."""
    d = q.shape[1]
    dq = _alloc_table(csr.n_rows, d, q.dtype, q.device) if out is None else out
    scaled = da if (scale_da_inplace and dst_scale is not None) else None
    if e_index is not None:
        if not want_de:
            if csr.n_rows:
                _edge_call("sirgcn_edge_bwd_q", csr, d, q.dtype, act, act_param, q, k, da, e, dq, None,
                           dst_scale, src_scale, scaled, e_index=e_index)
            return dq, None
        width = _pad_cols(d, q.dtype)
        n_units = csr.n_tiles + csr.n_chunks
        part = torch.zeros((max(n_units, 1), e.shape[0], width), dtype=torch.float32, device=q.device)
        if csr.n_rows:
            _edge_call("sirgcn_edge_bwd_q", csr, d, q.dtype, act, act_param, q, k, da, e, dq, None,
                       dst_scale, src_scale, scaled, e_index=e_index, de_partial=part)
        dtab = torch.empty((e.shape[0], d), dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            rc = _lib.lib().sirgcn_etable_grad(_lib.ptr(part), C.c_int64(n_units), C.c_int32(e.shape[0]), C.c_int32(width),
                                               C.c_int32(d), _lib.ptr(dtab), C.c_int64(d), _lib.stream_ptr(q.device))
        _lib.check(rc, "sirgcn_etable_grad")
        return dq, dtab
    de = _alloc_table(e.shape[0], d, q.dtype, q.device) if (want_de and e is not None) else None
    if csr.n_rows:
        _edge_call("sirgcn_edge_bwd_q", csr, d, q.dtype, act, act_param, q, k, da, e, dq, de,
                   dst_scale, src_scale, scaled)
    return dq, de


def edge_backward_k(csc, q, k, e, da, dst_scale, src_scale, act, act_param, out=None, e_index=None):
    d = k.shape[1]
    dk = _alloc_table(csc.n_rows, d, k.dtype, k.device) if out is None else out
    if csc.n_rows:
        _edge_call("sirgcn_edge_bwd_k", csc, d, k.dtype, act, act_param, q, k, da, e, dk, None,
                   dst_scale, src_scale, e_index=e_index)
    return dk


class EdgeAggregate(torch.autograd.Function):
    """A[u] = c_in[u] * sum_{e=(v->u)} c_out[v] * σ(Q[u] + K[v] + E[e])  with σ applied in registers."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, q, k, e, graph: Graph, agg_type, act, act_param):
        if not q.is_cuda:
            raise RuntimeError("SIR-GCN edge kernels need CUDA tensors (no CPU fallback)")
        if e is not None and graph.csr.eid is None:
            raise RuntimeError("edge features need a graph built with need_eid=True")
        if k.dtype != q.dtype or (e is not None and e.dtype != q.dtype):
            k = k.to(q.dtype)
            e = None if e is None else e.to(q.dtype)
        q, k = as_table(q.detach()), as_table(k.detach())
        e = None if e is None else as_table(e.detach())
        ds, ss = graph.scales(agg_type)
        out = edge_forward(graph.csr, q, k, e, ds, ss, act, act_param)
        ctx.save_for_backward(q, k, e)
        ctx.graph, ctx.agg_type, ctx.act, ctx.act_param = graph, agg_type, act, act_param
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        q, k, e = ctx.saved_tensors
        g = ctx.graph
        ds, ss = g.scales(ctx.agg_type)
        da = as_table(grad_out.to(q.dtype))
        need_q, need_k, need_e = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        dq = dk = de = None
        if need_q or (need_e and e is not None):
            dq, de = edge_backward_q(g.csr, q, k, e, da, ds, ss, ctx.act, ctx.act_param, need_e)
        if need_k:
            dk = edge_backward_k(g.csc, q, k, e, da, ds, ss, ctx.act, ctx.act_param)
        return dq, dk, de, None, None, None, None


class SIRLayerFunction(torch.autograd.Function):
    """One whole SIRConv/SIREConv call for the sum-like aggregators as ONE autograd node
    (SURVEY.md §8b): [Q|K] = H·W_qk^T + b  ->  fused edge stage  ->  out = A·W_R^T + b_R.
    Saves H, [Q|K], A (and the caller's projected edge term) — never an |E| x d tensor.  Backward
    writes dQ and dK straight into the two halves of one [N, 2·ld] buffer so the weight/input
    gradients of the concatenated projection are single GEMMs, and frees dA before dH is made."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, feat, w_qk, b_qk, e, w_r, b_r, graph, agg_type, act, act_param, d, recompute_qk=False,
                keep_q=None, keep_k=None, drop_scale=1.0, e_types=None):
        """keep_q / keep_k (torch.bool [N, d], or None): training-mode dropout of the two projections
        (conv.py:60-61), applied in place on the halves of the [Q|K] buffer; backward masks dQ / dK the same way.
        e_types (integer [E], or None): `e` is then a small TABLE [n_types, d] (the weight of an nn.Embedding edge
        term, zinc/model.py:12-15) and the edge term of edge i is row e_types[i] of it, looked up inside the kernels;
        the table's gradient is reduced inside the dQ walk.  No [E, d] tensor exists in either direction."""
        from . import gemm
        if not feat.is_cuda:
            raise RuntimeError("SIR-GCN kernels need CUDA tensors (no CPU fallback)")
        if e is not None and graph.csr.eid is None:
            raise RuntimeError("edge features need a graph built with need_eid=True")
        qk = gemm.linear_forward(feat, w_qk, b_qk)                 # [N, 2*ldp]
        ldp = qk.shape[1] // 2
        q, k = qk[:, :d], qk[:, ldp:ldp + d]
        q._sirgcn_padded = k._sirgcn_padded = True
        if keep_k is not None:
            mask_scale_(k, keep_k, drop_scale)
            mask_scale_(q, keep_q, drop_scale)
        if e is not None:
            e = as_table(e.detach().to(qk.dtype))
        ix_csr = ix_csc = None
        if e_types is not None:
            ix_csr, ix_csc = graph.edge_type_positions(e_types)
        ds, ss = graph.scales(agg_type)
        a = edge_forward(graph.csr, q, k, e, ds, ss, act, act_param, e_index=ix_csr)
        # recompute_qk: the [N, 2d] projection is not kept for backward but re-made from `feat` by one
        # more GEMM (≈5 % of the edge stage) — on the 50 M-node graph that is 25.6 GB per layer
        if recompute_qk:
            del q, k
            qk = None
        out = gemm.linear_forward(a, w_r, b_r)
        ctx.save_for_backward(feat, qk, e, a, w_qk, w_r, b_qk if recompute_qk else None, keep_q, keep_k)
        ctx.graph, ctx.agg_type, ctx.act, ctx.act_param, ctx.d = graph, agg_type, act, act_param, d
        ctx.drop_scale = drop_scale
        ctx.e_index = (ix_csr, ix_csc)
        ctx.has_bias = (b_qk is not None, b_r is not None)
        ctx.lean = bool(recompute_qk)      # the re-made [Q|K] belongs to backward alone: it may be overwritten
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, gout):
        from . import gemm
        feat, qk, e, a, w_qk, w_r, b_qk, keep_q, keep_k = ctx.saved_tensors
        g, d = ctx.graph, ctx.d
        remade = qk is None
        if remade:
            qk = gemm.linear_forward(feat, w_qk, b_qk)
        ldp = qk.shape[1] // 2
        q, k = qk[:, :d], qk[:, ldp:ldp + d]
        q._sirgcn_padded = k._sirgcn_padded = True
        if remade and keep_k is not None:
            mask_scale_(k, keep_k, ctx.drop_scale)
            mask_scale_(q, keep_q, ctx.drop_scale)
        need = ctx.needs_input_grad
        gout = gout.to(qk.dtype)
        gout = gout if gout.stride(-1) == 1 else gout.contiguous()
        dw_r = db_r = None
        if need[4]:         # the bias gradient rides along in the same pass over gout
            dw_r, db_r = gemm.linear_wgrad_bias(gout, a, w_r.dtype, need[5] and ctx.has_bias[1])
        elif need[5] and ctx.has_bias[1]:
            db_r = gemm.column_sum(gout, w_r.dtype)
        da = gemm.linear_dgrad(gout, w_r.to(qk.dtype), pad_to=_pad_cols(d, qk.dtype))[:, :d]   # [N, d]
        da._sirgcn_padded = True
        ds, ss = g.scales(ctx.agg_type)
        ix_csr, ix_csc = ctx.e_index
        if ctx.lean:
            # memory-lean order for tables of many GB (the 50 M-node graph): dQ goes to a buffer of its own, dK is
            # written IN PLACE over K (the CSC walk reads K[v] only as the own operand of row v, before it stores
            # dK[v]; long rows store in the finalize kernel, after every chunk has read K[v]), then dQ is copied over
            # Q — the [Q|K] buffer has become [dQ|dK] and no second [N, 2·ld] buffer ever existed: 12.8 GB less at
            # the peak for one extra 12.8 GB copy (≈ 4 ms of a 400 ms layer)
            dq_buf = _alloc_table(q.shape[0], d, qk.dtype, qk.device, zero=ldp != d)
            _, de = edge_backward_q(g.csr, q, k, e, da, ds, ss, ctx.act, ctx.act_param, need[3], out=dq_buf,
                                    scale_da_inplace=True, e_index=ix_csr)
            edge_backward_k(g.csc, q, k, e, da, None, ss, ctx.act, ctx.act_param, out=k, e_index=ix_csc)
            del da
            ldq = _pad_cols(d, qk.dtype)                    # whole 16-byte vectors, zero padding included
            copy_rows(qk[:, :ldq], dq_buf.as_strided((dq_buf.shape[0], ldq), (dq_buf.stride(0), 1)))
            del dq_buf
            dqk = qk
        else:
            dqk = (torch.empty if ldp == d else torch.zeros)(qk.shape, dtype=qk.dtype, device=qk.device)
            dq, dk = dqk[:, :d], dqk[:, ldp:ldp + d]
            # the dQ pass leaves dA scaled by the destination coefficient (in place: `da` is ours)
            _, de = edge_backward_q(g.csr, q, k, e, da, ds, ss, ctx.act, ctx.act_param, need[3], out=dq,
                                    scale_da_inplace=True, e_index=ix_csr)
            edge_backward_k(g.csc, q, k, e, da, None, ss, ctx.act, ctx.act_param, out=dk, e_index=ix_csc)
            del da
        del q, k, qk
        if keep_k is not None:                 # d(dropout): the same masks on the two gradient halves
            hq, hk = dqk[:, :d], dqk[:, ldp:ldp + d]
            hq._sirgcn_padded = hk._sirgcn_padded = True
            mask_scale_(hk, keep_k, ctx.drop_scale)
            mask_scale_(hq, keep_q, ctx.drop_scale)
        dw_qk = db_qk = None
        if need[1]:
            dw_qk, db_qk = gemm.linear_wgrad_bias(dqk, feat, w_qk.dtype, need[2] and ctx.has_bias[0])
        elif need[2] and ctx.has_bias[0]:
            db_qk = gemm.column_sum(dqk, w_qk.dtype)
        dfeat = gemm.linear_dgrad(dqk, w_qk.to(dqk.dtype)).to(feat.dtype) if need[0] else None
        if de is not None and ix_csr is not None:
            de = de.to(w_qk.dtype)                  # the table's gradient, in the parameter's dtype
        return dfeat, dw_qk, db_qk, de, dw_r, db_r, None, None, None, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# split path
# ----------------------------------------------------------------------------------------------
def _gather_add(num_pos, parts, d, dtype, device):
    """parts: list of (table, selector) with table [*, d] (unit column stride)."""
    z = torch.empty((num_pos, d), dtype=dtype, device=device)
    if num_pos == 0:
        return z
    parts = parts + [(None, None)] * (3 - len(parts))
    args = []
    for _, sel in parts:
        args.append(_lib.ptr(sel))
    for t, _ in parts:
        args += [_lib.ptr(t), C.c_int64(0 if t is None else t.stride(0))]
    with torch.cuda.device(device):
        rc = _lib.lib().sirgcn_gather_add(C.c_int64(num_pos), *args, _lib.ptr(z), C.c_int64(d), C.c_int32(d),
                                          C.c_int32(_lib.DTYPE_CODE[dtype]), _lib.stream_ptr(device))
    _lib.check(rc, "sirgcn_gather_add")
    return z


def _segment_sum(rows: CompressedRows, perm, m, d, dst_scale, src_scale):
    out = torch.empty((rows.n_rows, d), dtype=m.dtype, device=m.device)
    if rows.n_rows == 0:
        return out
    with torch.cuda.device(m.device):
        rc = _lib.lib().sirgcn_segment_sum(
            C.c_int32(rows.n_rows), _lib.ptr(rows.indptr), _lib.ptr(rows.idx), _lib.ptr(perm),
            _lib.ptr(m), C.c_int64(m.stride(0)), _lib.ptr(out), C.c_int64(d), C.c_int32(d),
            C.c_int32(_lib.DTYPE_CODE[m.dtype]), _lib.ptr(dst_scale), _lib.ptr(src_scale), _lib.stream_ptr(m.device))
    _lib.check(rc, "sirgcn_segment_sum")
    return out


def _rowmajor(t):
    return t if (t.dim() == 2 and t.stride(1) == 1) else t.contiguous()


class GatherAdd(torch.autograd.Function):
    """z[p] = Q[dst(p)] + K[src(p)] + E[eid(p)] for every CSR position p (any operand may be None)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, q, k, e, graph: Graph):
        ref = next(t for t in (q, k, e) if t is not None)
        if not ref.is_cuda:
            raise RuntimeError("SIR-GCN kernels need CUDA tensors (no CPU fallback)")
        dtype, d = ref.dtype, ref.shape[1]
        parts = []
        if q is not None:
            parts.append((_rowmajor(q.detach().to(dtype)), graph.pos_dst()))
        if k is not None:
            parts.append((_rowmajor(k.detach().to(dtype)), graph.csr.idx))
        if e is not None:
            parts.append((_rowmajor(e.detach().to(dtype)), graph.pos_eid()))
        ctx.graph = graph
        ctx.has = (q is not None, k is not None, e is not None)
        return _gather_add(graph.num_edges(), parts, d, dtype, ref.device)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dz):
        g = ctx.graph
        dz = _rowmajor(dz)
        d = dz.shape[1]
        dq = dk = de = None
        if ctx.has[0] and ctx.needs_input_grad[0]:
            dq = _segment_sum(g.csr, None, dz, d, None, None)
        if ctx.has[1] and ctx.needs_input_grad[1]:
            dk = _segment_sum(g.csc, g.csc_pos(), dz, d, None, None)
        if ctx.has[2] and ctx.needs_input_grad[2]:
            de = torch.empty_like(dz)
            de[g.pos_eid().long()] = dz
        return dq, dk, de, None


class SegmentReduce(torch.autograd.Function):
    """DGL builtin reducers over CSR-ordered messages m [E, d] (conv.py:41,63): sum / mean / sym
    (coefficients folded in) or max / min with argmax bookkeeping (0 for empty rows)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, m, graph: Graph, agg_type):
        if not m.is_cuda:
            raise RuntimeError("SIR-GCN kernels need CUDA tensors (no CPU fallback)")
        m = _rowmajor(m.detach())
        d = m.shape[1]
        ctx.graph, ctx.agg_type = graph, agg_type
        rows = graph.csr
        if agg_type in ("max", "min"):
            out = torch.empty((rows.n_rows, d), dtype=m.dtype, device=m.device)
            arg = torch.empty((rows.n_rows, d), dtype=torch.int32, device=m.device)
            if rows.n_rows:
                with torch.cuda.device(m.device):
                    rc = _lib.lib().sirgcn_segment_minmax(
                        C.c_int32(rows.n_rows), _lib.ptr(rows.indptr), _lib.ptr(m), C.c_int64(m.stride(0)),
                        _lib.ptr(out), C.c_int64(d), _lib.ptr(arg), C.c_int64(d), C.c_int32(d),
                        C.c_int32(_lib.DTYPE_CODE[m.dtype]), C.c_int32(agg_type == "min"), _lib.stream_ptr(m.device))
                _lib.check(rc, "sirgcn_segment_minmax")
            ctx.save_for_backward(arg)
            return out
        ds, ss = graph.scales(agg_type)
        return _segment_sum(rows, None, m, d, ds, ss)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        g, agg = ctx.graph, ctx.agg_type
        dout = _rowmajor(dout)
        d, E = dout.shape[1], g.num_edges()
        if agg in ("max", "min"):
            (arg,) = ctx.saved_tensors
            dm = torch.empty((E, d), dtype=dout.dtype, device=dout.device)
            if E:
                with torch.cuda.device(dout.device):
                    rc = _lib.lib().sirgcn_segment_minmax_bwd(
                        C.c_int64(E), _lib.ptr(g.pos_dst()), _lib.ptr(dout), C.c_int64(dout.stride(0)),
                        _lib.ptr(arg), C.c_int64(d), _lib.ptr(dm), C.c_int64(d), C.c_int32(d),
                        C.c_int32(_lib.DTYPE_CODE[dout.dtype]), _lib.stream_ptr(dout.device))
                _lib.check(rc, "sirgcn_segment_minmax_bwd")
            return dm, None, None
        dm = _gather_add(E, [(dout, g.pos_dst())], d, dout.dtype, dout.device)
        ds, ss = g.scales(agg)
        if ds is not None:
            coef = ds[g.pos_dst().long()]
            if ss is not None:
                coef = coef * ss[g.csr.idx.long()]
            dm = dm * coef.to(dm.dtype).unsqueeze(1)
        return dm, None, None
