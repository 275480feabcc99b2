"""dgl.transforms.DropEdge (test stand-in, see dgl/__init__.py): /root/reference/models/utils.py:96 subclasses it."""
import torch


class DropEdge:
    def __init__(self, p=0.5):
        self.p = p

    def __call__(self, g):
        from . import DGLGraph
        if self.p == 0:
            keep = torch.ones(g.num_edges(), dtype=torch.bool, device=g.device)
        else:
            keep = torch.rand(g.num_edges(), device=g.device) >= self.p
        out = DGLGraph(g._src[keep], g._dst[keep], g.num_nodes())
        out._batch_num_nodes = g._batch_num_nodes
        for k, v in g.ndata.items():
            out.ndata[k] = v
        for k, v in g.edata.items():
            out.edata[k] = v[keep]
        return out
