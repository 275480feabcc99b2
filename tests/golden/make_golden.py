"""Generates tests/golden/sirconv_golden.pt — frozen input/output vectors of the ORACLE
(oracle/sirconv_ref.py) under fixed seeds.

The reference layer itself cannot be imported here (it needs DGL, which is not installed), so these
vectors pin the oracle against silent drift and give the GPU tests a fixture that does not depend on
the oracle's code at run time.  If a DGL-equipped machine is available, run with --check-dgl to diff
the unmodified reference layer against the same fixtures.

    python tests/golden/make_golden.py            # rewrite the fixture
"""
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sirconv_ref import RefGraph, RefSIRConv, RefSIREConv, csr_csc_ref  # noqa: E402

ACTS = {"relu": nn.ReLU, "leaky": lambda: nn.LeakyReLU(0.2), "gelu": nn.GELU, "identity": nn.Identity}


def make_case(seed, n, e, d_in, d, d_out, act, agg, edge_dim=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    if n > 3:                      # isolated destinations and sources
        dst[dst == 1] = 0
        src[src == 2] = 0
    torch.manual_seed(seed)
    if edge_dim:
        layer = RefSIREConv(d_in, edge_dim, d, d_out, ACTS[act](), agg_type=agg)
    else:
        layer = RefSIRConv(d_in, d, d_out, ACTS[act](), agg_type=agg)
    feat = torch.randn(n, d_in, generator=g, requires_grad=True)
    efeat = torch.randn(e, edge_dim, generator=g, requires_grad=True) if edge_dim else None
    out = layer(RefGraph(src, dst, n), feat, efeat)
    gout = torch.randn(out.shape, generator=g)
    params = list(layer.parameters())
    grads = torch.autograd.grad(out, [feat] + ([efeat] if edge_dim else []) + params, gout)
    case = {
        "meta": dict(seed=seed, n=n, e=e, d_in=d_in, d=d, d_out=d_out, act=act, agg=agg, edge_dim=edge_dim),
        "src": src.int(), "dst": dst.int(), "feat": feat.detach(), "efeat": None if efeat is None else efeat.detach(),
        "state": {k: v.detach().clone() for k, v in layer.state_dict().items()},
        "out": out.detach(), "gout": gout, "dfeat": grads[0],
        "defeat": grads[1] if edge_dim else None,
        "dparams": {k: gr for (k, _), gr in zip(layer.named_parameters(), grads[(2 if edge_dim else 1):])},
        "csr": csr_csc_ref(src, dst, n),
    }
    return case


def main():
    cases = []
    seed = 100
    for agg in ("sum", "mean", "sym", "max"):
        for act in ("relu", "leaky", "gelu"):
            cases.append(make_case(seed, 37, 160, 12, 16, 10, act, agg)); seed += 1
    cases.append(make_case(seed, 37, 160, 12, 20, 10, "leaky", "sym", edge_dim=3)); seed += 1   # d=20: padded rows
    cases.append(make_case(seed, 29, 90, 8, 75, 6, "leaky", "sum", edge_dim=4)); seed += 1      # published odd size
    cases.append(make_case(seed, 1, 0, 4, 8, 4, "relu", "sum")); seed += 1                      # no edges
    cases.append(make_case(seed, 5, 40, 4, 8, 4, "identity", "mean")); seed += 1                # dense multigraph
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sirconv_golden.pt")
    torch.save(cases, out)
    print(f"wrote {len(cases)} cases to {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    main()
