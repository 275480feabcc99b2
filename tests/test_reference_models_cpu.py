"""BASELINE.json configs[0] — "synthetic-datasets/dictionary-lookup SIR-GCN train step on CPU via DGL (reference
plumbing, no GPU)" — and its sibling hetero-edge-count: the reference's OWN model files, imported unmodified (their DGL
imports served by tests/fake_dgl), against the same models assembled from the restated oracle layers.

This pins the oracle one level above the layer: embeddings -> SIRConv stack with the non-elementwise
σ = Sequential(ReLU, Linear, ReLU) (dictionary-lookup/model.py:17) / ReLU (hetero-edge-count/model.py:17) -> dropout ->
classifier / regression + SumPooling, forward, loss, backward and one AdamW step.  CPU only; skipped where the reference
tree is not mounted (the GPU box).  The CUDA side of the same shapes is tests/test_configs_gpu.py.
"""
import importlib.util
import os
import sys
import types

import pytest
import torch
from torch import nn

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models", "conv.py")),
                                reason="the reference tree is not on this box")
FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_dgl")

from oracle.sirconv_ref import RefGraph, RefSIRConv  # noqa: E402


def _load_reference_model(rel_path):
    """exec <reference>/<rel_path> unmodified; `models` resolves to the REFERENCE's package while it loads (this repo
    has a drop-in `models` package of its own), `dgl` to the stand-in"""
    if FAKE not in sys.path:
        sys.path.insert(0, FAKE)
    import dgl  # noqa: F401
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    try:
        spec = importlib.util.spec_from_file_location("_ref_model_" + rel_path.replace("/", "_").replace("-", "_"),
                                                      os.path.join(REF, rel_path))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.SIRConv.__module__ == "models.conv" and "reference" in sys.modules["models.conv"].__file__
        return mod, dgl
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _dictionary_batch(dgl, num_nodes, graphs, seed):
    """synthetic-datasets/dictionary-lookup/data.py:21-39: complete bipartite value -> key graphs, (key, value) ids"""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(graphs):
        key, val = torch.arange(0, num_nodes), torch.arange(num_nodes, 2 * num_nodes)
        gr = dgl.graph((val.repeat_interleave(num_nodes), key.repeat(num_nodes)), num_nodes=2 * num_nodes)
        perm = torch.randperm(num_nodes, generator=g)
        gr.ndata["feat"] = torch.cat([torch.stack([key, torch.full_like(key, num_nodes)], 1), torch.stack([key, perm], 1)])
        gr.ndata["mask"] = torch.cat([torch.ones(num_nodes, dtype=torch.bool), torch.zeros(num_nodes, dtype=torch.bool)])
        out.append(gr)
    return dgl.batch(out)


class _OracleDictionaryModel(nn.Module):
    """dictionary-lookup/model.py:11-35 with the restated layer"""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.key_embedding = nn.Embedding(input_dim + 1, hidden_dim)
        self.val_embedding = nn.Embedding(input_dim + 1, hidden_dim)
        self.activation = nn.Sequential(nn.ReLU(inplace=True), nn.Linear(hidden_dim, hidden_dim), nn.ReLU(inplace=True))
        self.convs = nn.ModuleList([RefSIRConv(hidden_dim, hidden_dim, hidden_dim, self.activation) for _ in range(num_layers)])
        self.drop = nn.Dropout(0)
        self.classifier = nn.Linear(hidden_dim, output_dim, bias=False)

    def forward(self, graph, feats):
        h = self.key_embedding(feats[:, 0]) + self.val_embedding(feats[:, 1])
        for conv in self.convs:
            h = self.drop(conv(graph, h))
        return self.classifier(h)


@pytest.mark.parametrize("num_layers", [1, 2])
def test_dictionary_lookup_train_step_matches_reference_model(num_layers):
    ref_mod, dgl = _load_reference_model("synthetic-datasets/dictionary-lookup/model.py")
    nodes, hidden = 10, 40                                      # train.py defaults of BASELINE configs[0]
    torch.manual_seed(0)
    ref = ref_mod.SIRModel(nodes, hidden, nodes, num_layers=num_layers).double()
    mine = _OracleDictionaryModel(nodes, hidden, nodes, num_layers).double()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    graphs = _dictionary_batch(dgl, nodes, graphs=16, seed=1)
    feats, mask = graphs.ndata.pop("feat"), graphs.ndata.pop("mask")
    labels = feats[:, 1].to(torch.int64)
    src, dst = graphs.edges()
    rg = RefGraph(src, dst, graphs.num_nodes())
    loss_fn = nn.CrossEntropyLoss()
    opt_a = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    opt_b = torch.optim.AdamW(mine.parameters(), lr=1e-3)
    for _ in range(2):                                          # two steps: the second sees updated weights
        la = loss_fn(ref(graphs, feats)[mask], labels[~mask])
        lb = loss_fn(mine(rg, feats)[mask], labels[~mask])
        torch.testing.assert_close(lb, la, rtol=1e-12, atol=1e-12)
        opt_a.zero_grad(), opt_b.zero_grad()
        la.backward(), lb.backward()
        for (name, pa), pb in zip(ref.named_parameters(), mine.parameters()):
            torch.testing.assert_close(pb.grad, pa.grad, rtol=1e-10, atol=1e-12, msg=name)
        opt_a.step(), opt_b.step()
    assert len(graphs.ndata) == 0                               # the layer left the caller's graph untouched


def test_hetero_edge_count_model_matches_reference_model():
    ref_mod, dgl = _load_reference_model("synthetic-datasets/hetero-edge-count/model.py")
    classes, hidden = 5, 16
    torch.manual_seed(0)
    ref = ref_mod.SIRModel(classes, hidden, 1, num_layers=2).double()
    gen = torch.Generator().manual_seed(3)
    gs = []
    for _ in range(6):                                          # hetero-edge-count/data.py:27-31 (dgl.rand_graph)
        n = int(torch.randint(2, 9, (1,), generator=gen))
        e = int(torch.randint(n * n // 4, n * n + 1, (1,), generator=gen))
        g = dgl.rand_graph(n, e, generator=gen)
        g.ndata["label"] = torch.randint(0, classes, (n,), generator=gen)
        gs.append(g)
    graphs = dgl.batch(gs)
    labels = graphs.ndata.pop("label")
    out_ref = ref(graphs, labels)
    src, dst = graphs.edges()
    rg = RefGraph(src, dst, graphs.num_nodes())
    convs = [RefSIRConv(hidden, hidden, hidden, nn.ReLU(inplace=True)).double() for _ in range(2)]
    for c, r in zip(convs, ref.convs):
        c.load_state_dict(r.state_dict())
    h = ref.embedding(labels)
    for c in convs:
        h = c(rg, h)
    h = ref.regression(h)
    sizes = graphs.batch_num_nodes()
    gid = torch.repeat_interleave(torch.arange(sizes.numel()), sizes)
    out = torch.zeros(sizes.numel(), 1, dtype=h.dtype).index_add_(0, gid, h)
    torch.testing.assert_close(out, out_ref, rtol=1e-12, atol=1e-12)
    ga = torch.autograd.grad(out_ref.sum(), list(ref.convs.parameters()))
    gb = torch.autograd.grad(out.sum(), [p for c in convs for p in c.parameters()])
    for a, b in zip(ga, gb):
        torch.testing.assert_close(b, a, rtol=1e-10, atol=1e-12)


def test_reference_dropedge_wrapper_runs_on_the_stand_in():
    """models/utils.py:96-102 (the per-layer DropEdge of the benchmark models): edge features follow the kept edges"""
    if FAKE not in sys.path:
        sys.path.insert(0, FAKE)
    import dgl
    import importlib
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    try:
        ru = importlib.import_module("models.utils")
        g = dgl.rand_graph(30, 400, generator=torch.Generator().manual_seed(0))
        ef = torch.arange(400.0).unsqueeze(1)
        torch.manual_seed(0)
        g2, ef2 = ru.DropEdge(0.5)(g, ef)
        assert 100 < g2.num_edges() < 300 and ef2.shape[0] == g2.num_edges()
        kept = ef2.squeeze(1).long()
        assert torch.equal(g2.edges()[0], g.edges()[0][kept]) and torch.equal(g2.edges()[1], g.edges()[1][kept])
        g3, ef3 = ru.DropEdge(0.0)(g, ef)
        assert g3.num_edges() == 400 and torch.equal(ef3, ef) and "efeats_" not in g.edata
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
