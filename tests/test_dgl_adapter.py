"""a DGLGraph handed to the layers (`as_graph`, sir-gcn_b200/graph.py).

DGL is not installable here; tests/fake_dgl models the slice of its graph API the adapter touches
(`edges(form='uv', order='eid')`, `num_nodes()`, int64 ids, `ndata` / `edata` frames) — the same stand-in that lets the
unmodified reference layer run for the golden fixtures.
"""
import gc
import os
import sys

import pytest
import torch
from torch import nn

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_dgl")
if FAKE not in sys.path:
    sys.path.insert(0, FAKE)
import dgl  # noqa: E402  (the stand-in)

import sirgcn_b200  # noqa: E402,F401
from sirgcn_b200 import Graph, SIRConv, SIREConv, as_graph  # noqa: E402
from sirgcn_b200 import graph as graph_mod  # noqa: E402


def _coo(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)


def test_cpu_dglgraph_is_rejected_loudly():
    """the product path has no CPU fallback: a DGLGraph that lives on the host must raise, not compute"""
    src, dst = _coo(10, 30, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        as_graph(dgl.graph((src, dst), num_nodes=10))
    with pytest.raises(TypeError):
        as_graph(object())


@pytest.mark.gpu
def test_dglgraph_int64_edges_convert_once_and_are_not_mutated():
    dev = "cuda:0"
    n, e = 300, 2500
    src, dst = _coo(n, e, 1)
    dg = dgl.graph((src, dst), num_nodes=n).to(dev)
    assert dg.edges()[0].dtype == torch.int64 and dg.edges()[0].is_cuda
    dg.ndata["keep_me"] = torch.arange(n, device=dev)
    dg.edata["w"] = torch.ones(e, device=dev)
    torch.manual_seed(0)
    layer = SIREConv(16, 3, 32, 8, nn.LeakyReLU(0.2), agg_type="sym").to(dev)
    x = torch.randn(n, 16, device=dev, requires_grad=True)
    ef = torch.randn(e, 3, device=dev)
    out_dgl = layer(dg, x, ef)
    (dx_dgl,) = torch.autograd.grad(out_dgl, x, torch.ones_like(out_dgl))
    own = Graph(src.to(dev), dst.to(dev), n)
    out_own = layer(own, x, ef)
    (dx_own,) = torch.autograd.grad(out_own, x, torch.ones_like(out_own))
    assert torch.equal(out_dgl, out_own) and torch.equal(dx_dgl, dx_own)          # same kernels, same structures
    # converted ONCE per DGLGraph object: the second call finds the cached Graph
    first = as_graph(dg)
    assert as_graph(dg) is first
    for a, b in ((first.csr.indptr, own.csr.indptr), (first.csr.idx, own.csr.idx), (first.csr.eid, own.csr.eid),
                 (first.csc.indptr, own.csc.indptr), (first.csc.idx, own.csc.idx), (first.csc.eid, own.csc.eid)):
        assert a.dtype == torch.int32 and torch.equal(a, b)
    # conv.py:50 local_scope: the caller's graph is untouched
    assert set(dg.ndata) == {"keep_me"} and set(dg.edata) == {"w"}
    assert torch.equal(dg.edges()[0].cpu(), src) and torch.equal(dg.edges()[1].cpu(), dst)
    # the cache is weak: dropping the DGLGraph drops its converted structures
    before = len(graph_mod._dgl_cache)
    del dg, first
    gc.collect()
    assert len(graph_mod._dgl_cache) == before - 1


@pytest.mark.gpu
def test_unmodified_reference_layer_on_gpu_matches_cuda_layer():
    """when the reference tree is present (this container's image on a GPU box it is not): the reference layer itself,
    on the GPU through the stand-in, against the CUDA layer on the very same DGLGraph object"""
    if not os.path.exists("/root/reference/models/conv.py"):
        pytest.skip("the reference tree is not on this box")
    from tests.golden.make_golden import load_reference
    ref, _ = load_reference()
    dev = "cuda:0"
    n, e = 120, 900
    src, dst = _coo(n, e, 2)
    dg = dgl.graph((src, dst), num_nodes=n).to(dev)
    torch.manual_seed(0)
    a = ref.SIRConv(8, 16, 8, nn.ReLU(), agg_type="mean").to(dev)
    b = SIRConv(8, 16, 8, nn.ReLU(), agg_type="mean").to(dev)
    b.load_state_dict(a.state_dict())
    x = torch.randn(n, 8, device=dev)
    torch.testing.assert_close(b(dg, x), a(dg, x), rtol=1e-5, atol=1e-5)
