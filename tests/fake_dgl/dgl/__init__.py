"""A ~150-line stand-in for the slice of DGL 2.1 that /root/reference/models/conv.py touches — TEST INFRASTRUCTURE.

Purpose: let the UNMODIFIED reference layer (`models/conv.py`, which does `from dgl import function as fn` and
`from dgl.utils import expand_as_pair` at import time) execute on CPU, so that (a) the golden fixtures under
tests/golden/ are produced by the reference's own code and (b) the restated oracle (oracle/sirconv_ref.py) is
checked against it.  DGL itself (dgl==2.1.0, requirements.txt:1) is not installable here.

The documented semantics of the eight symbols the layer uses are modelled (first block); a few more (second block)
exist so that the reference's own MODEL files (synthetic-datasets/*/model.py, which import models.utils and with it
dgl.transforms.DropEdge, and use dgl.nn.SumPooling) import and run unmodified for the CPU plumbing of BASELINE.json
configs[0]:

    dgl.graph((src, dst), num_nodes=)      homogeneous multigraph, edge id = position in the COO list
    g.num_nodes() / g.num_edges() / g.device / g.edges(form, order) / g.in_degrees() / g.out_degrees()
    g.ndata / g.edata                      dict-like frames
    g.local_scope()                        frames are restored on exit (the caller's graph is not mutated)
    g.update_all(message_udf, reducer)     message UDF sees an EdgeBatch with .src / .dst / .data views gathered in
                                           edge-id order; builtin reducers fn.sum / fn.mean / fn.max / fn.min reduce
                                           over the in-edges of every node; nodes without in-edges get ZEROS
                                           (DGL's documented behaviour for every builtin reducer, max/min included)
    dgl.function.{sum,mean,max,min}(msg, out)
    dgl.utils.expand_as_pair(feat, g)      (feat, feat) on a homogeneous graph, tuples pass through

    dgl.batch(graphs)                      block-diagonal union: node ids offset, frames concatenated, batch_size /
                                           batch_num_nodes() kept
    dgl.rand_graph(n, e)                   random multigraph (self loops allowed), as the hetero-edge-count data uses
    dgl.transforms.DropEdge(p)(g)          every edge removed independently with probability p; edata follows
    dgl.nn.SumPooling()(g, feat)           per-graph sums of a batched graph

The reducers are written differently from the oracle on purpose (dense incidence products and per-node Python loops
instead of index_add_ / scatter_reduce_), so that the two do not share an implementation.
"""
from __future__ import annotations

import contextlib

import torch

from . import function, nn, transforms, utils  # noqa: F401

__version__ = "2.1.0+fake"


class _Frame(dict):
    """ndata / edata: a dict of tensors whose first dimension is checked on assignment (as DGL does)"""

    def __init__(self, rows):
        super().__init__()
        self._rows = rows

    def __setitem__(self, key, value):
        if not torch.is_tensor(value) or value.shape[0] != self._rows:
            raise ValueError(f"Expect number of features to match number of rows ({self._rows}), got "
                             f"{tuple(value.shape) if torch.is_tensor(value) else type(value)}")
        super().__setitem__(key, value)


class EdgeBatch:
    def __init__(self, g):
        self._g = g
        self.src = {k: v.index_select(0, g._src) for k, v in g.ndata.items()}
        self.dst = {k: v.index_select(0, g._dst) for k, v in g.ndata.items()}
        self.data = dict(g.edata)

    def __len__(self):
        return self._g.num_edges()


class DGLGraph:
    is_block = False

    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src).to(torch.int64)      # DGL's default idtype
        self._dst = torch.as_tensor(dst).to(torch.int64)
        self._n = int(num_nodes)
        if self._src.numel() and (int(torch.max(self._src.max(), self._dst.max())) >= self._n
                                  or int(torch.min(self._src.min(), self._dst.min())) < 0):
            raise ValueError("node ids out of range")
        self.ndata = _Frame(self._n)
        self.edata = _Frame(int(self._src.numel()))
        self._batch_num_nodes = None            # set by dgl.batch

    # --- structure queries ---------------------------------------------------------------------------------
    @property
    def device(self):
        return self._src.device

    @property
    def batch_size(self):
        return 1 if self._batch_num_nodes is None else int(self._batch_num_nodes.numel())

    def batch_num_nodes(self):
        return torch.tensor([self._n]) if self._batch_num_nodes is None else self._batch_num_nodes

    def to(self, device):
        g = DGLGraph(self._src.to(device), self._dst.to(device), self._n)
        g._batch_num_nodes = None if self._batch_num_nodes is None else self._batch_num_nodes.to(device)
        for k, v in self.ndata.items():
            g.ndata[k] = v.to(device)
        for k, v in self.edata.items():
            g.edata[k] = v.to(device)
        return g

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def edges(self, form="uv", order="eid"):
        if form != "uv" or order != "eid":
            raise NotImplementedError("fake dgl: edges(form='uv', order='eid') only")
        return self._src, self._dst

    def in_degrees(self):
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src, minlength=self._n)

    # --- frames --------------------------------------------------------------------------------------------
    @contextlib.contextmanager
    def local_scope(self):
        saved_n, saved_e = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            dict.clear(self.ndata), dict.update(self.ndata, saved_n)
            dict.clear(self.edata), dict.update(self.edata, saved_e)

    # --- message passing -----------------------------------------------------------------------------------
    def update_all(self, message_func, reduce_func):
        if not callable(message_func) or not isinstance(reduce_func, function.BuiltinReducer):
            raise NotImplementedError("fake dgl: update_all(UDF message, builtin reducer) only")
        msgs = message_func(EdgeBatch(self))
        m = msgs[reduce_func.msg_field]
        if m.shape[0] != self.num_edges():
            raise ValueError("message must have one row per edge")
        self.ndata[reduce_func.out_field] = reduce_func.reduce(m, self._dst, self._n)


def batch(graphs):
    """block-diagonal union of homogeneous graphs (node ids offset by the sizes of the graphs before)"""
    graphs = list(graphs)
    sizes = torch.tensor([g.num_nodes() for g in graphs])
    offs = torch.cumsum(sizes, 0) - sizes
    src = torch.cat([g._src + int(o) for g, o in zip(graphs, offs)]) if graphs else torch.empty(0, dtype=torch.int64)
    dst = torch.cat([g._dst + int(o) for g, o in zip(graphs, offs)]) if graphs else torch.empty(0, dtype=torch.int64)
    out = DGLGraph(src, dst, int(sizes.sum()))
    for key in (graphs[0].ndata.keys() if graphs else ()):
        out.ndata[key] = torch.cat([g.ndata[key] for g in graphs])
    for key in (graphs[0].edata.keys() if graphs else ()):
        out.edata[key] = torch.cat([g.edata[key] for g in graphs])
    out._batch_num_nodes = sizes
    return out


def rand_graph(num_nodes, num_edges, generator=None):
    src = torch.randint(0, num_nodes, (num_edges,), generator=generator)
    dst = torch.randint(0, num_nodes, (num_edges,), generator=generator)
    return DGLGraph(src, dst, num_nodes)


def graph(data, num_nodes=None, idtype=None, device=None):
    src, dst = data
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if num_nodes is None:
        num_nodes = int(max(src.max(), dst.max())) + 1 if src.numel() else 0
    return DGLGraph(src, dst, num_nodes)
