"""1-D destination-row partition of one large graph over the GPUs of a node (one process per GPU,
torch.distributed / NCCL over NVLink).  New capability — the reference is single-process — whose
correctness oracle is "G-rank result == 1-rank result" (SURVEY.md §8e).

Rank r owns node rows [r*n_pad, min(N, (r+1)*n_pad)), n_pad = ceil(N/G): its slice of H, Q, K, A, the
in-CSR rows of those destinations (column ids stay GLOBAL source ids) and the out-CSC rows of those
sources (row ids stay GLOBAL destination ids).  Per layer and direction there is exactly one exchange:

    forward   K_full  = all_gather(K_loc)                 -> edge_fwd   over the local CSR rows
    backward  dQ_loc  = edge_bwd_q over the local CSR rows (re-uses K_full)
              Q_full, dA_full = all_gather(Q_loc), all_gather(dA_loc)
              dK_loc  = edge_bwd_k over the local CSC rows
    weights   replicated; dW summed with all_reduce

Because gathered tables are laid out [G*n_pad, d], a global node id indexes them directly.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import function as F_
from . import gemm
from .graph import CompressedRows, Graph


class CudaEdgeBackend:
    """the product backend: C-ABI CUDA kernels"""
    forward = staticmethod(F_.edge_forward)
    backward_q = staticmethod(F_.edge_backward_q)
    backward_k = staticmethod(F_.edge_backward_k)


class RowPartition:
    """Local slice of a graph for rank `rank` of `world` (see module docstring)."""

    def __init__(self, num_nodes, rank, world, csr_local, csc_local, in_norm, out_norm, inv_in_deg, group=None):
        self.num_nodes_, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        self.n_pad = (self.num_nodes_ + world - 1) // world
        self.lo = min(self.num_nodes_, rank * self.n_pad)
        self.hi = min(self.num_nodes_, self.lo + self.n_pad)
        self.csr, self.csc = csr_local, csc_local
        # per-node coefficients, padded to G*n_pad so that gathered-table indices are valid
        self.in_norm, self.out_norm, self.inv_in_deg = in_norm, out_norm, inv_in_deg
        self.num_local_edges = csr_local.num_pos

    @property
    def n_local(self):
        return self.hi - self.lo

    @classmethod
    def from_graph(cls, graph: Graph, rank, world, group=None):
        """slice an already converted (replicated) Graph; the caller may drop `graph` afterwards"""
        n = graph.num_nodes()
        n_pad = (n + world - 1) // world
        lo = min(n, rank * n_pad)
        hi = min(n, lo + n_pad)
        pad = lambda t: torch.cat([t, t.new_ones(world * n_pad - n)]) if world * n_pad > n else t
        return cls(n, rank, world, graph.csr.slice_rows(lo, hi), graph.csc.slice_rows(lo, hi),
                   pad(graph.in_norm), pad(graph.out_norm), pad(graph.inv_in_deg), group)

    def scales_rows(self, agg_type):
        """(dst_scale, src_scale) for the CSR walks: dst = local row, src = global id"""
        if agg_type == "sym":
            return self.in_norm[self.lo:self.lo + max(self.n_local, 1)], self.out_norm
        if agg_type == "mean":
            return self.inv_in_deg[self.lo:self.lo + max(self.n_local, 1)], None
        return None, None

    def scales_cols(self, agg_type):
        """(dst_scale, src_scale) for the CSC walk: dst = global id, src = local row"""
        if agg_type == "sym":
            return self.in_norm, self.out_norm[self.lo:self.lo + max(self.n_local, 1)]
        if agg_type == "mean":
            return self.inv_in_deg, None
        return None, None

    def all_gather_rows(self, local):
        """[n_local, d] -> [G*n_pad, d] (rows of rank r at r*n_pad ...), padded rows are zero"""
        d = local.shape[1]
        if local.shape[0] != self.n_pad or not local.is_contiguous():
            buf = local.new_zeros((self.n_pad, d))
            buf[: local.shape[0]].copy_(local)
            local = buf
        full = local.new_empty((self.world * self.n_pad, d))
        if self.world == 1:
            full.copy_(local)
        else:
            dist.all_gather_into_tensor(full, local, group=self.group)
        return full

    def all_reduce_(self, t):
        if self.world > 1 and t is not None:
            dist.all_reduce(t, group=self.group)
        return t


class PartitionedSIRLayerFunction(torch.autograd.Function):
    """SIRLayerFunction for a row-partitioned graph: same arithmetic per row, one all-gather per
    direction, weight gradients all-reduced (so every rank ends with the full-graph gradient)."""

    @staticmethod
    def forward(ctx, feat_loc, w_qk, b_qk, w_r, b_r, part: RowPartition, agg_type, act, act_param, d, backend):
        qk = gemm.linear_forward(feat_loc, w_qk, b_qk)
        ldp = qk.shape[1] // 2
        q, k = qk[:, :d], qk[:, ldp:ldp + d]
        q._sirgcn_padded = True
        k_full = part.all_gather_rows(k)
        ds, ss = part.scales_rows(agg_type)
        a = backend.forward(part.csr, q, k_full, None, ds, ss, act, act_param)
        out = gemm.linear_forward(a, w_r, b_r)
        ctx.save_for_backward(feat_loc, qk, k_full, a, w_qk, w_r)
        ctx.part, ctx.agg_type, ctx.act, ctx.act_param, ctx.d, ctx.backend = part, agg_type, act, act_param, d, backend
        ctx.has_bias = (b_qk is not None, b_r is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        feat, qk, k_full, a, w_qk, w_r = ctx.saved_tensors
        part, d, be = ctx.part, ctx.d, ctx.backend
        ldp = qk.shape[1] // 2
        q, k = qk[:, :d], qk[:, ldp:ldp + d]
        q._sirgcn_padded = k._sirgcn_padded = True
        need = ctx.needs_input_grad
        gout = gout.to(qk.dtype)
        gout = gout if gout.stride(-1) == 1 else gout.contiguous()
        dw_r = part.all_reduce_(gemm.linear_wgrad(gout, a, w_r.dtype)) if need[3] else None
        db_r = part.all_reduce_(gout.sum(0).to(w_r.dtype)) if (need[4] and ctx.has_bias[1]) else None
        da = gemm.linear_dgrad(gout, w_r.to(qk.dtype), pad_to=F_._pad_cols(d, qk.dtype))[:, :d]
        da._sirgcn_padded = True
        dqk = (torch.empty if ldp == d else torch.zeros)(qk.shape, dtype=qk.dtype, device=qk.device)
        dq, dk = dqk[:, :d], dqk[:, ldp:ldp + d]
        ds, ss = part.scales_rows(ctx.agg_type)
        be.backward_q(part.csr, q, k_full, None, da, ds, ss, ctx.act, ctx.act_param, False, out=dq)
        del k_full
        q_full, da_full = part.all_gather_rows(q), part.all_gather_rows(da)
        del da
        ds, ss = part.scales_cols(ctx.agg_type)
        be.backward_k(part.csc, q_full, k, None, da_full, ds, ss, ctx.act, ctx.act_param, out=dk)
        del q_full, da_full
        dw_qk = part.all_reduce_(gemm.linear_wgrad(dqk, feat, w_qk.dtype)) if need[1] else None
        db_qk = part.all_reduce_(dqk.sum(0).to(w_qk.dtype)) if (need[2] and ctx.has_bias[0]) else None
        dfeat = gemm.linear_dgrad(dqk, w_qk.to(qk.dtype)).to(feat.dtype) if need[0] else None
        return dfeat, dw_qk, db_qk, dw_r, db_r, None, None, None, None, None, None


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
    w, b, d, _ = layer._cat_qk_weights(feat_loc.dtype)
    lr = layer.linear_relation
    return PartitionedSIRLayerFunction.apply(feat_loc, w, b, lr.weight, lr.bias, part, layer._agg_type,
                                             known[0], known[1], d, backend)
