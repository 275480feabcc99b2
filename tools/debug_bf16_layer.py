"""debug: bf16 whole-layer path vs the fp64 oracle, tensor by tensor (smoke()'s second half, opened up)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
import sirgcn_b200
from sirgcn_b200 import Graph, SIRConv
from oracle.sirconv_ref import RefGraph, RefSIRConv

dev = "cuda:0"
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-20))

def run(n, e, hub, agg, act, d_in, d, d_out, dtype, seed=0):
    torch.manual_seed(seed)
    src, dst = torch.randint(0, n, (e,)), torch.randint(0, n, (e,))
    if hub:
        dst[:hub] = 3
    ref = RefSIRConv(d_in, d, d_out, act(), agg_type=agg)
    lay = SIRConv(d_in, d, d_out, act(), agg_type=agg).to(dev)
    lay.load_state_dict(ref.state_dict())
    x, go = torch.randn(n, d_in), torch.randn(n, d_out)
    cast = (lambda t: t.to(dtype).double()) if dtype != torch.float32 else (lambda t: t.double())
    xr = cast(x).requires_grad_(True)
    ref = ref.double()
    o_ref = ref(RefGraph(src, dst, n), xr)
    gr = torch.autograd.grad(o_ref, [xr] + list(ref.parameters()), cast(go))
    g = Graph(src.to(dev), dst.to(dev), n, long_threshold=256)
    xb = x.to(dev).to(dtype).requires_grad_(True)
    o = lay(g, xb)
    gg = torch.autograd.grad(o, [xb] + list(lay.parameters()), go.to(dev).to(dtype))
    names = ["dfeat"] + [k for k, _ in lay.named_parameters()]
    print(f"n={n} e={e} hub={hub} {agg} {act.__name__} {d_in}->{d}->{d_out} {dtype}: out {rel(o, o_ref):.2e} | " +
          " ".join(f"{nm} {rel(a, b):.2e}" for nm, a, b in zip(names, gg, gr)), flush=True)

print("SIRGCN_GEMM =", os.environ.get("SIRGCN_GEMM"))
for dtype in (torch.bfloat16, torch.float32):
    run(200, 1500, 700, "mean", nn.ReLU, 64, 128, 64, dtype)
    run(200, 1500, 0, "mean", nn.ReLU, 64, 128, 64, dtype)
    run(200, 1500, 700, "sum", nn.ReLU, 64, 128, 64, dtype)
    run(200, 1500, 700, "mean", nn.GELU, 64, 128, 64, dtype)
    run(256, 1500, 700, "mean", nn.ReLU, 64, 128, 64, dtype)
    run(200, 1500, 700, "mean", nn.ReLU, 128, 128, 128, dtype)
    run(2000, 15000, 700, "mean", nn.ReLU, 64, 128, 64, dtype)
