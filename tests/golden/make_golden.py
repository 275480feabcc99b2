"""Generates tests/golden/sirconv_golden.pt by EXECUTING THE UNMODIFIED REFERENCE LAYERS.

    python tests/golden/make_golden.py            # needs /root/reference (this container); rewrites the fixture

`/root/reference/models/conv.py` is loaded as-is (importlib, no edits).  Its two imports of DGL
(`from dgl import function as fn`, `from dgl.utils import expand_as_pair`, conv.py:3-4) are served by the small
stand-in under tests/fake_dgl/ (DGL 2.1.0 itself is not installable here), which models the documented semantics of
the eight DGL symbols the layer touches.  Everything else — the order of the projections and dropouts, the degree
clamps and norms, the message function, which aggregators apply W_R per edge, bias placement — is the reference's
own code running.

Each case stores fp32-valued inputs / weights and the reference's outputs and gradients evaluated in fp64 (the layer
is `.double()`-ed, so the stored results are exact to ~1e-15 for those fp32 inputs).  Consumers:
  * tests/test_oracle_pins.py    the restated oracle must reproduce every case to 1e-12 (CPU);
  * tests/test_gpu_parity.py     the CUDA layers must reproduce every case to 1e-5 (fp32 tables) on the B200.
The GPU box has no /root/reference: only the committed .pt travels.
"""
import importlib.util
import os
import sys

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE_CONV = "/root/reference/models/conv.py"
FAKE_DGL = os.path.join(ROOT, "tests", "fake_dgl")

ACTS = {"relu": nn.ReLU, "leaky": lambda: nn.LeakyReLU(0.2), "gelu": nn.GELU, "identity": nn.Identity}
AGGS = ("sum", "mean", "sym", "max", "min")


def load_reference():
    """(reference conv module, fake dgl module) — the reference file is executed unmodified"""
    if not os.path.exists(REFERENCE_CONV):
        raise FileNotFoundError(REFERENCE_CONV)
    if FAKE_DGL not in sys.path:
        sys.path.insert(0, FAKE_DGL)
    import dgl
    if not dgl.__version__.endswith("+fake"):        # a real DGL would be even better; say which one ran
        print(f"note: a real dgl {dgl.__version__} is importable and is being used")
    spec = importlib.util.spec_from_file_location("_reference_models_conv", REFERENCE_CONV)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, dgl


def message_mlp(d_cat, d_out):
    """the callable g of the *Base layers (a small MLP, as synthetic-datasets/dictionary-lookup/model.py:17 uses)"""
    return nn.Sequential(nn.Linear(d_cat, 2 * d_out), nn.ReLU(), nn.Linear(2 * d_out, d_out))


def build_layer(ns, meta):
    """instantiate class `meta['cls']` from namespace `ns` (the reference module, the oracle, or the CUDA package)"""
    cls, agg = meta["cls"], meta["agg"]
    if cls in ("SIRConv", "SIREConv"):
        act = ACTS[meta["act"]]()
        if cls == "SIRConv":
            return ns["SIRConv"](meta["d_in"], meta["d"], meta["d_out"], act, agg_type=agg,
                                 inner_bias=meta["inner_bias"], outer_bias=meta["outer_bias"])
        return ns["SIREConv"](meta["d_in"], meta["edge_dim"], meta["d"], meta["d_out"], act, agg_type=agg,
                              inner_bias=meta["inner_bias"], outer_bias=meta["outer_bias"])
    d_cat = 2 * meta["d_in"] + (meta["edge_dim"] if cls == "SIREConvBase" else 0)
    return ns[cls](message_mlp(d_cat, meta["d_out"]), agg_type=agg)


def make_case(ref, dgl, seed, cls, n, e, d_in, d, d_out, act, agg, edge_dim=0, inner_bias=True, outer_bias=True):
    gen = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=gen)
    dst = torch.randint(0, n, (e,), generator=gen)
    if n > 3:                      # a destination without in-edges and a source without out-edges
        dst[dst == 1] = 0
        src[src == 2] = 0
    meta = dict(seed=seed, cls=cls, n=n, e=e, d_in=d_in, d=d, d_out=d_out, act=act, agg=agg, edge_dim=edge_dim,
                inner_bias=inner_bias, outer_bias=outer_bias)
    torch.manual_seed(seed)
    ns = {k: getattr(ref, k) for k in ("SIRConv", "SIREConv", "SIRConvBase", "SIREConvBase")}
    layer = build_layer(ns, meta)                      # fp32 default init = fp32-representable weights
    state32 = {k: v.detach().clone() for k, v in layer.state_dict().items()}
    layer = layer.double()
    feat = torch.randn(n, d_in, generator=gen).double().requires_grad_(True)
    has_e = cls in ("SIREConv", "SIREConvBase")
    efeat = torch.randn(e, edge_dim, generator=gen).double().requires_grad_(True) if has_e else None
    g = dgl.graph((src, dst), num_nodes=n)
    out = layer(g, feat, efeat) if has_e else layer(g, feat)
    assert len(g.ndata) == 0 and len(g.edata) == 0, "the reference must leave the caller's graph untouched"
    gout = torch.randn(out.shape, generator=gen).double()
    params = list(layer.named_parameters())
    grads = torch.autograd.grad(out, [feat] + ([efeat] if has_e else []) + [p for _, p in params], gout,
                                allow_unused=True)
    k0 = 2 if has_e else 1
    return {
        "meta": meta, "src": src.int(), "dst": dst.int(),
        "feat": feat.detach().float(), "efeat": None if efeat is None else efeat.detach().float(),
        "state": state32, "out": out.detach(), "gout": gout.float(), "dfeat": grads[0],
        "defeat": grads[1] if has_e else None,
        "dparams": {name: (torch.zeros_like(p) if gr is None else gr) for (name, p), gr in zip(params, grads[k0:])},
    }


def main():
    ref, dgl = load_reference()
    sys.path.insert(0, ROOT)
    from oracle.sirconv_ref import csr_csc_ref
    cases, seed = [], 100
    for agg in AGGS:
        for act in ("relu", "leaky", "gelu"):
            cases.append(make_case(ref, dgl, seed, "SIRConv", 37, 160, 12, 16, 10, act, agg)); seed += 1
    for agg in AGGS:
        cases.append(make_case(ref, dgl, seed, "SIREConv", 37, 160, 12, 20, 10, "leaky", agg, edge_dim=3)); seed += 1
        cases.append(make_case(ref, dgl, seed, "SIRConvBase", 23, 90, 6, 0, 5, "", agg)); seed += 1
        cases.append(make_case(ref, dgl, seed, "SIREConvBase", 23, 90, 6, 0, 5, "", agg, edge_dim=2)); seed += 1
    cases.append(make_case(ref, dgl, seed, "SIREConv", 29, 90, 8, 75, 6, "leaky", "sum", edge_dim=4)); seed += 1   # odd size
    cases.append(make_case(ref, dgl, seed, "SIRConv", 1, 0, 4, 8, 4, "relu", "sum")); seed += 1                   # no edges
    cases.append(make_case(ref, dgl, seed, "SIRConv", 5, 40, 4, 8, 4, "identity", "mean")); seed += 1             # dense multigraph
    cases.append(make_case(ref, dgl, seed, "SIRConv", 40, 300, 8, 12, 7, "leaky", "sym", inner_bias=False,
                           outer_bias=False)); seed += 1
    cases.append(make_case(ref, dgl, seed, "SIRConv", 64, 2000, 16, 32, 16, "relu", "mean")); seed += 1           # hubs
    for c in cases:
        c["csr"] = csr_csc_ref(c["src"], c["dst"], c["meta"]["n"])
    out = os.path.join(HERE, "sirconv_golden.pt")
    torch.save(cases, out)
    print(f"wrote {len(cases)} cases produced by {REFERENCE_CONV} to {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    main()
