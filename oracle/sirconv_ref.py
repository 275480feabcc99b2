"""CPU oracle for the SIR-GCN convolution — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The shipped path
(``sir-gcn_b200/``, ``models/``) never does: it fails loudly if the CUDA library
is missing.

PINNED ON THE REFERENCE'S OWN CODE (not on DGL's): the reference layer (/root/reference/models/conv.py) imports
``dgl`` at module import time (conv.py:3-4); DGL 2.1.0 (requirements.txt:1) is not installable here and the
reference ships no tests / golden vectors.  tests/golden/make_golden.py therefore executes the UNMODIFIED conv.py
with its eight DGL symbols served by the stand-in under tests/fake_dgl/ and commits the results
(tests/golden/sirconv_golden.pt: 4 classes x 5 aggregators, fp64); tests/test_oracle_pins.py requires this
restatement to reproduce them — outputs and all gradients — to 1e-12, and repeats the comparison live on fresh
cases wherever /root/reference exists.  What remains unpinned is DGL's own arithmetic behind those eight symbols
(documented semantics only).  This file restates the layer's arithmetic op-for-op with the ATen operations DGL
lowers ``update_all(UDF message, builtin reduce)`` to:

* edge-UDF gathers  -> ``index_select``          (conv.py:45  edges.src[...]/edges.dst[...])
* message           -> elementwise add/σ/mul      (conv.py:43-47, 109-113)
* ``fn.sum``        -> ``index_add_`` over dst    (conv.py:63, 130)
* ``fn.mean``       -> sum / clamp(in_degree, 1)  (DGL's gspmm 'mean' = sum then divide)
* ``fn.max``        -> ``scatter_reduce_('amax', include_self=False)``, empty rows = 0
* degrees           -> ``bincount`` + ``clamp(min=1)``   (conv.py:51-52)

Further, independent pins (tests/test_oracle_pins.py): fp64 ``gradcheck``, the
hetero-edge-count exact-count identity (synthetic-datasets/hetero-edge-count/data.py:21),
the dictionary-lookup isolated-destination identity (dictionary-lookup/data.py:27-31),
algebraic identities (sym on a regular graph, mean = sum/deg, edge permutation
invariance, edge duplication) and frozen golden vectors (tests/golden/).
"""
from __future__ import annotations

import torch
from torch import nn


class RefGraph:
    """Minimal stand-in for the slice of DGLGraph the layer consumes
    (conv.py:50-52,55,63): COO edge list src->dst over ``num_nodes`` nodes."""

    def __init__(self, src, dst, num_nodes):
        self.src = torch.as_tensor(src, dtype=torch.int64)
        self.dst = torch.as_tensor(dst, dtype=torch.int64)
        self.n = int(num_nodes)

    def num_nodes(self):
        return self.n

    def num_edges(self):
        return int(self.src.numel())

    def in_degrees(self):
        return torch.bincount(self.dst, minlength=self.n)

    def out_degrees(self):
        return torch.bincount(self.src, minlength=self.n)


def _reduce(graph: RefGraph, msg: torch.Tensor, how: str) -> torch.Tensor:
    """DGL builtin reducers over the in-edges of every node (conv.py:41,63)."""
    out = msg.new_zeros((graph.n,) + tuple(msg.shape[1:]))
    if how in ("sum", "sym"):
        out.index_add_(0, graph.dst, msg)
    elif how == "mean":
        out.index_add_(0, graph.dst, msg)
        deg = graph.in_degrees().clamp(min=1).to(msg.dtype)
        out = out / deg.reshape((-1,) + (1,) * (msg.dim() - 1))
    elif how in ("max", "min"):
        idx = graph.dst.reshape((-1,) + (1,) * (msg.dim() - 1)).expand_as(msg)
        red = "amax" if how == "max" else "amin"
        out.scatter_reduce_(0, idx, msg, red, include_self=False)
    else:
        raise AttributeError(f"module 'dgl.function' has no attribute '{how}'")
    return out


def _norms(graph: RefGraph, agg: str, like: torch.Tensor):
    """conv.py:51-57 — clamped degrees, ^-1/2 only for 'sym'.  The reference computes them with `.float()`
    (conv.py:51-52) whatever the dtype of the features: they are fp32 values that the per-edge product then promotes
    (found by running the unmodified layer in fp64 against this restatement, tests/test_oracle_pins.py)."""
    shape = (graph.n,) + (1,) * (like.dim() - 1)
    if agg == "sym":
        i = torch.pow(graph.in_degrees().float().clamp(min=1), -0.5)
        o = torch.pow(graph.out_degrees().float().clamp(min=1), -0.5)
    else:
        i = torch.ones(graph.n)
        o = torch.ones(graph.n)
    return i.reshape(shape), o.reshape(shape)


def _store(t, dtype):
    """round to the table dtype and come back (straight-through for autograd); identity when dtype is None"""
    if dtype is None or t is None:
        return t
    return t + (t.detach().to(dtype).to(t.dtype) - t.detach())


class RefSIRConv(nn.Module):
    """Restatement of SIRConv (conv.py:7-67): same constructor, sub-module names
    and state_dict, so weights can be moved to/from the CUDA layer verbatim.

    ``storage_dtype`` (None by default = the plain layer) models the reference run in 16-bit (``model.bfloat16()`` /
    autocast): the projections eq / ek / e, the aggregate and the output are ROUNDED to that dtype where the 16-bit
    run stores them, while this oracle keeps evaluating in fp64.  With a discontinuous σ' (ReLU) the sign of
    z = eq + ek is decided by the rounded values, so a 16-bit run can only be compared with an oracle that rounds
    at the same places."""
    storage_dtype = None

    def __init__(self, input_dim, hidden_dim, output_dim, activation, dropout=0,
                 inner_bias=True, outer_bias=True, agg_type="sum"):
        super().__init__()
        self.activation = activation
        self.dropout = nn.Dropout(dropout)
        self.linear_query = nn.Linear(input_dim, hidden_dim, bias=inner_bias)
        self.linear_key = nn.Linear(input_dim, hidden_dim, bias=False)
        self.linear_relation = nn.Linear(hidden_dim, output_dim, bias=outer_bias)
        self._agg_type = agg_type

    def _edge_term(self, graph, efeat):
        return None

    def forward(self, graph: RefGraph, feat, efeat=None):
        agg = self._agg_type
        in_norm, out_norm = _norms(graph, agg, feat)
        sd = self.storage_dtype
        ek = _store(self.dropout(self.linear_key(feat)), sd)          # conv.py:60 (K first)
        eq = _store(self.dropout(self.linear_query(feat)), sd)        # conv.py:61
        z = eq.index_select(0, graph.dst) + ek.index_select(0, graph.src)
        e = _store(self._edge_term(graph, efeat), sd)
        if e is not None:
            z = z + e                                     # conv.py:111
        if agg in ("sum", "mean", "sym"):
            m = out_norm.index_select(0, graph.src) * in_norm.index_select(0, graph.dst) * self.activation(z)
            return _store(self.linear_relation(_store(_reduce(graph, m, agg), sd)), sd)   # conv.py:63-65
        m = self.linear_relation(self.activation(z))      # conv.py:47
        return _store(_reduce(graph, m, agg), sd)


class RefSIREConv(RefSIRConv):
    """Restatement of SIREConv (conv.py:70-134)."""

    def __init__(self, input_dim, edge_dim, hidden_dim, output_dim, activation, dropout=0,
                 inner_bias=True, outer_bias=True, agg_type="sum"):
        super().__init__(input_dim, hidden_dim, output_dim, activation, dropout,
                         inner_bias, outer_bias, agg_type)
        self.linear_edge = nn.Linear(edge_dim, hidden_dim, bias=False)

    def _edge_term(self, graph, efeat):
        return self.dropout(self.linear_edge(efeat))      # conv.py:128

    def forward(self, graph, nfeat, efeat):
        return super().forward(graph, nfeat, efeat)


class RefSIRConvBase(nn.Module):
    """Restatement of SIRConvBase (conv.py:137-177): g([h_dst ‖ h_src]) per edge."""

    def __init__(self, message_func, agg_type="sum"):
        super().__init__()
        self._agg_type = agg_type
        self._message_func = message_func

    def forward(self, graph, feat, efeat=None):
        in_norm, out_norm = _norms(graph, self._agg_type, feat)
        parts = [feat.index_select(0, graph.dst), feat.index_select(0, graph.src)]
        if efeat is not None:
            parts.append(efeat)                           # conv.py:199
        m = self._message_func(torch.cat(parts, dim=-1))
        m = out_norm.index_select(0, graph.src) * in_norm.index_select(0, graph.dst) * m
        return _reduce(graph, m, self._agg_type)


class RefSIREConvBase(RefSIRConvBase):
    """Restatement of SIREConvBase (conv.py:180-221)."""

    def forward(self, graph, nfeat, efeat):
        return super().forward(graph, nfeat, efeat)


# ---------------------------------------------------------------------------
# index construction oracle ("index/CSR construction bit-exact")
# ---------------------------------------------------------------------------
def csr_csc_ref(src, dst, num_nodes):
    """Destination-sorted CSR and source-sorted CSC by STABLE sort (ties keep edge-id
    order).  Returns int32 tensors (indptr_in, col_src, eid_in, indptr_out, row_dst,
    eid_out) plus float32 (in_norm, out_norm) = clamp(deg,1)^-1/2 (conv.py:51-57)."""
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    n = int(num_nodes)
    _, eid_in = torch.sort(dst, stable=True)
    _, eid_out = torch.sort(src, stable=True)
    indeg = torch.bincount(dst, minlength=n)
    outdeg = torch.bincount(src, minlength=n)
    z = torch.zeros(1, dtype=torch.int64)
    indptr_in = torch.cat([z, indeg.cumsum(0)])
    indptr_out = torch.cat([z, outdeg.cumsum(0)])
    i32 = lambda t: t.to(torch.int32)
    in_norm = indeg.clamp(min=1).to(torch.float32).pow(-0.5)
    out_norm = outdeg.clamp(min=1).to(torch.float32).pow(-0.5)
    return (i32(indptr_in), i32(src[eid_in]), i32(eid_in),
            i32(indptr_out), i32(dst[eid_out]), i32(eid_out), in_norm, out_norm)
