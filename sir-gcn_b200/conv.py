"""B200-native SIR-GCN layers — drop-in for /root/reference/models/conv.py.

Same constructors, sub-module names (=> identical ``state_dict``) and ``forward(graph, feat[, efeat])``
as the reference ``SIRConv`` (conv.py:7-67), ``SIREConv`` (:70-134), ``SIRConvBase`` (:137-177) and
``SIREConvBase`` (:180-221).  ``graph`` may be this package's ``Graph`` or a ``DGLGraph``.

Data flow of one call (sum / mean / sym, σ ∈ {ReLU, LeakyReLU, GELU(erf), Identity}):

    [Q | K] = feat · [W_Q ; W_K]^T + [b_Q | 0]        one concatenated GEMM      (conv.py:60-61)
    A       = EdgeAggregate(graph, Q, K, E)            fused CUDA edge stage      (conv.py:43-47,:63)
    out     = A · W_R^T + b_R                          GEMM                       (conv.py:65)

Any other σ (an arbitrary callable, e.g. Sequential(ReLU, Linear, ReLU)), ``agg_type`` max/min and
the *Base layers go through the split path GatherAdd -> callable -> SegmentReduce.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .function import EdgeAggregate, GatherAdd, SegmentReduce, SIRLayerFunction, _pad_cols, draw_keep_mask
from .gemm import linear as _linear
from .graph import as_graph

_SUM_LIKE = ("sum", "mean", "sym")
_AGG_TYPES = ("sum", "mean", "sym", "max", "min")   # what getattr(dgl.function, agg_type) admits (conv.py:41)


def classify_activation(act):
    """(code, param) when σ is an elementwise activation the fused kernels implement, else None."""
    if isinstance(act, nn.ReLU) or act in (torch.relu, F.relu):
        return _lib.ACT_RELU, 0.0
    if isinstance(act, nn.LeakyReLU):
        return _lib.ACT_LEAKY_RELU, float(act.negative_slope)
    if isinstance(act, nn.GELU) and act.approximate == "none":
        return _lib.ACT_GELU, 0.0
    if isinstance(act, nn.Identity):
        return _lib.ACT_IDENTITY, 0.0
    return None


def _check_agg(agg_type):
    if agg_type not in _AGG_TYPES:
        raise AttributeError(f"module 'dgl.function' has no attribute '{agg_type}'")


class SIRConv(nn.Module):
    r"""h*_u = Σ_{v∈N(u)} W_R σ(W_Q h_u + W_K h_v)   (reference: models/conv.py:7-67)"""

    def __init__(self, input_dim, hidden_dim, output_dim, activation, dropout=0, inner_bias=True,
                 outer_bias=True, agg_type="sum"):
        super().__init__()
        _check_agg(agg_type)
        self.activation = activation
        self.dropout = nn.Dropout(dropout)
        self.linear_query = nn.Linear(input_dim, hidden_dim, bias=inner_bias)
        self.linear_key = nn.Linear(input_dim, hidden_dim, bias=False)
        self.linear_relation = nn.Linear(hidden_dim, output_dim, bias=outer_bias)
        self._agg_type = agg_type
        self._agg_func = "sum" if agg_type == "sym" else agg_type   # name of the DGL builtin (conv.py:41)
        self.recompute_qk = None    # None = auto (see forward); not part of the reference surface

    # -- projections -------------------------------------------------------------------------
    def _plain(self):
        lq, lk, lr = self.linear_query, self.linear_key, self.linear_relation
        return (type(lq) is nn.Linear and type(lk) is nn.Linear and type(lr) is nn.Linear
                and lk.bias is None and lq.weight.shape == lk.weight.shape)

    def _cat_qk_weights(self, dtype):
        """[W_Q ; W_K] (+ zero rows so that each half is a whole number of 16-B vectors for odd
        hidden sizes such as the published 75 / 95) and [b_Q | 0]."""
        wq, wk, bq = self.linear_query.weight, self.linear_key.weight, self.linear_query.bias
        d = wq.shape[0]
        ldp = _pad_cols(d, dtype)
        bq = bq if bq is not None else wq.new_zeros(d)
        if ldp != d:
            zw, zb = wq.new_zeros(ldp - d, wq.shape[1]), wq.new_zeros(ldp - d)
            return torch.cat([wq, zw, wk, zw]), torch.cat([bq, zb, wq.new_zeros(ldp)]), d, ldp
        return torch.cat([wq, wk]), torch.cat([bq, wq.new_zeros(d)]), d, ldp

    def _project_qk(self, feat):
        """K then Q (dropout RNG order of conv.py:60-61) from ONE concatenated GEMM."""
        if not self._plain():
            k = self.dropout(self.linear_key(feat))
            return self.dropout(self.linear_query(feat)), k
        dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else feat.dtype
        w, b, d, ldp = self._cat_qk_weights(dt)
        qk = _linear(feat, w, b)
        q, k = qk[:, :d], qk[:, ldp:ldp + d]
        if self.training and self.dropout.p > 0:
            # nn.Dropout draws its mask per memory layout: on contiguous [N, d] tensors the K and Q masks
            # are bit-identical to the reference's two separate projections (K drawn first, conv.py:60-61)
            k = self.dropout(k.contiguous())
            q = self.dropout(q.contiguous())
        else:
            q._sirgcn_padded = k._sirgcn_padded = True
        return q, k

    def _edge_term(self, graph, efeat):
        return None

    def _edge_term_or_table(self, graph, efeat, dropping):
        """(e, e_types) for the whole-layer node: (projected edge term [E, d], None), or — when the edge term is an
        nn.Embedding over at most MAX_ETYPES edge types and no dropout applies to it — (the embedding TABLE, the
        integer edge types): the kernels then look the rows up themselves and reduce the table's gradient, and no
        [E, d] tensor exists in either direction."""
        return self._edge_term(graph, efeat), None

    def forward(self, graph, feat, efeat=None):
        g = as_graph(graph)
        if feat.dim() < 2:
            raise ValueError("feat must be [N, ..., input_dim]")
        if feat.shape[0] != g.num_nodes():
            raise ValueError(f"feat has {feat.shape[0]} rows but the graph has {g.num_nodes()} nodes")
        n, inner = feat.shape[0], tuple(feat.shape[1:-1])   # conv.py:55 tolerates [N, ..., d_in]
        agg = self._agg_type
        known = classify_activation(self.activation)
        dropping = self.training and self.dropout.p > 0
        # the fused kernels hold one row in at most 128 16-byte vectors (2048 B: 512 fp32 / 1024 16-bit columns); wider
        # hidden sizes take the split path below, which has no width limit (the reference accepts any size)
        dt_ = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else feat.dtype
        d_hid = self.linear_query.weight.shape[0] if hasattr(self.linear_query, "weight") else 0
        fits = 0 < d_hid and _pad_cols(d_hid, dt_) * torch.empty((), dtype=dt_).element_size() <= 2048
        if agg in _SUM_LIKE and known is not None and not inner and fits and self._plain():
            # one autograd node for the whole layer, in training mode too: the dropout masks of the two projections
            # are drawn here, K first then Q (conv.py:60-61; the edge term's mask is the third draw, conv.py:128), and
            # applied inside the node on the halves of the [Q|K] buffer
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else feat.dtype
            w, b, d, _ = self._cat_qk_weights(dt)
            keep_q = keep_k = None
            scale = 1.0
            if dropping:
                p = float(self.dropout.p)
                keep_k = draw_keep_mask(n, d, dt, feat.device, p)
                keep_q = draw_keep_mask(n, d, dt, feat.device, p)
                scale = 0.0 if p >= 1 else 1.0 / (1.0 - p)
            e, e_types = self._edge_term_or_table(g, efeat, dropping)
            lr = self.linear_relation
            recompute = self.recompute_qk
            if recompute is None:     # auto: do not keep a projection larger than 4 GiB for backward
                recompute = 2 * n * _pad_cols(d, dt) * torch.empty((), dtype=dt).element_size() > (1 << 32)
            return SIRLayerFunction.apply(feat, w, b, e, lr.weight, lr.bias, g, agg, known[0], known[1], d,
                                          bool(recompute), keep_q, keep_k, scale, e_types)
        q, k = self._project_qk(feat.reshape(-1, feat.shape[-1]))
        e = self._edge_term(g, efeat)
        if agg in _SUM_LIKE and known is not None and not inner and fits:
            a = EdgeAggregate.apply(q, k, e, g, agg, known[0], known[1])
            if type(self.linear_relation) is nn.Linear:
                return _linear(a, self.linear_relation.weight, self.linear_relation.bias)
            return self.linear_relation(a)
        # split path: materialise z in CSR order, run the callable, reduce.  Every op of the edge
        # stage acts per trailing column, so inner dims are folded into the column axis.
        E = g.num_edges()
        z = GatherAdd.apply(q.reshape(n, -1), k.reshape(n, -1), None if e is None else e.reshape(E, -1), g)
        s = self.activation(z.reshape((E,) + inner + (-1,)))
        if agg in _SUM_LIKE:
            a = SegmentReduce.apply(s.reshape(E, -1), g, agg)
            return self.linear_relation(a.reshape((n,) + inner + (-1,)))
        m = self.linear_relation(s)                     # conv.py:47 — W_R per edge, then max/min
        out = SegmentReduce.apply(m.reshape(E, -1), g, agg)
        return out.reshape((n,) + inner + (-1,))


class SIREConv(SIRConv):
    r"""h*_u = Σ_v W_R σ(W_Q h_u + W_E h_uv + W_K h_v)   (reference: models/conv.py:70-134).
    ``linear_edge`` may be replaced after construction (e.g. nn.Embedding,
    benchmark-datasets/zinc/model.py:12-15); whatever it is, it is called on ``efeat``."""

    def __init__(self, input_dim, edge_dim, hidden_dim, output_dim, activation, dropout=0, inner_bias=True,
                 outer_bias=True, agg_type="sum"):
        super().__init__(input_dim, hidden_dim, output_dim, activation, dropout, inner_bias, outer_bias, agg_type)
        self.linear_edge = nn.Linear(edge_dim, hidden_dim, bias=False)

    def _edge_term(self, graph, efeat):
        if efeat.shape[0] != graph.num_edges():
            raise ValueError(f"efeat has {efeat.shape[0]} rows but the graph has {graph.num_edges()} edges")
        return self.dropout(self.linear_edge(efeat))    # conv.py:128 (third dropout draw)

    def _edge_term_or_table(self, graph, efeat, dropping):
        le = self.linear_edge
        if (type(le) is nn.Embedding and not dropping and le.num_embeddings <= _lib.MAX_ETYPES and le.padding_idx is None
                and le.max_norm is None and not le.scale_grad_by_freq and not le.sparse
                and torch.is_tensor(efeat) and efeat.dim() == 1 and not efeat.is_floating_point()
                and not efeat.dtype == torch.bool):
            if efeat.shape[0] != graph.num_edges():
                raise ValueError(f"efeat has {efeat.shape[0]} rows but the graph has {graph.num_edges()} edges")
            return le.weight, efeat
        return self._edge_term(graph, efeat), None

    def forward(self, graph, nfeat, efeat):
        return super().forward(graph, nfeat, efeat)


class SIRConvBase(nn.Module):
    r"""h*_u = Σ_v g([h_u ‖ h_v])   (reference: models/conv.py:137-177)"""

    def __init__(self, message_func, agg_type="sum"):
        super().__init__()
        _check_agg(agg_type)
        self._agg_type = agg_type
        self._agg_func = "sum" if agg_type == "sym" else agg_type   # name of the DGL builtin (conv.py:154)
        self._message_func = message_func

    def forward(self, graph, feat, efeat=None):
        g = as_graph(graph)
        if feat.dim() < 2 or feat.shape[0] != g.num_nodes():
            raise ValueError(f"feat must be [N={g.num_nodes()}, ..., d], got {tuple(feat.shape)}")
        n, E = g.num_nodes(), g.num_edges()
        # conv.py:159-166 tolerates [N, ..., d]: the gathers act per trailing column, so inner dims are folded into
        # the column axis for the kernels and restored before the concatenation along the LAST axis (conv.py:157)
        shape_e = (E,) + tuple(feat.shape[1:])
        f2 = feat.reshape(n, -1)
        parts = [GatherAdd.apply(f2, None, None, g).reshape(shape_e), GatherAdd.apply(None, f2, None, g).reshape(shape_e)]
        if efeat is not None:
            if efeat.shape[0] != E:
                raise ValueError(f"efeat has {efeat.shape[0]} rows but the graph has {E} edges")
            parts.append(GatherAdd.apply(None, None, efeat.reshape(E, -1), g).reshape(efeat.shape))
        m = self._message_func(torch.cat(parts, dim=-1))
        out = SegmentReduce.apply(m.reshape(E, -1), g, self._agg_type)
        return out.reshape((n,) + tuple(m.shape[1:]))


class SIREConvBase(SIRConvBase):
    r"""h*_u = Σ_v g([h_u ‖ h_uv ‖ h_v])   (reference: models/conv.py:180-221; note the reference
    concatenates in the order dst, src, edge — conv.py:199)"""

    def forward(self, graph, nfeat, efeat):
        return super().forward(graph, nfeat, efeat)
