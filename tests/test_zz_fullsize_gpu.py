"""BASELINE.json configs[4] at its FULL size (50 M nodes / 2 B edges, d = 128, bf16) on one B200, through
size-independent properties — the CPU oracle cannot run at this size.  Named test_zz_* so that it runs last.

With Q = 0 and every K row equal to one positive vector c, σ = ReLU and mean aggregation:
  forward   A[u]  = c for every destination with in-edges, 0 for the others (DGL's zero fill)          -> exact
  backward  dQ[u] = dA[u] · [in_deg(u) > 0]   (Σ_e dA[u]/deg over deg edges: exact when dA is a power of two)
            dK[v] = Σ_{e: v->u} dA[u] / in_deg(u): a checksum over ALL rows (Σ_v dK[v] = Σ_u in_deg(u)·dA[u]/in_deg(u))
            plus sampled rows (and the largest ones) against an fp64 gather-sum
Every row of the forward and dQ walks is checked exactly, including the hub rows that go through the chunk schedule.
"""
import pytest
import torch

import sirgcn_b200  # noqa: F401
from sirgcn_b200 import Graph, _lib, synth
from sirgcn_b200 import function as F_

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_powerlaw_2b_edges_exact_properties():
    free, total = torch.cuda.mem_get_info(0)
    if free < 130 * 2**30:
        pytest.skip(f"needs ~130 GiB of free HBM, {free / 2**30:.0f} GiB available")
    n, e, d = 50_000_000, 2_000_000_000, 128
    src, dst, n = synth.powerlaw_hashed(n, e, alpha=2.3, seed=0, device=DEV)
    g = Graph(src, dst, n, need_eid=False, keep_coo=False)
    del src, dst
    assert g.csr.num_pos == e and g.csc.num_pos == e and g.csr.n_chunks > 0
    indeg, outdeg = g.in_degrees(), g.out_degrees()
    assert int(indeg.sum()) == e and int(outdeg.sum()) == e
    dt = torch.bfloat16
    c = (torch.arange(d, device=DEV, dtype=torch.float32) % 7 + 1).to(dt)          # small integers: exact in bf16
    q = torch.zeros((n, d), dtype=dt, device=DEV)
    k = c.expand(n, d).contiguous()
    ds, ss = g.scales("mean")
    a = F_.edge_forward(g.csr, q, k, None, ds, ss, _lib.ACT_RELU, 0.0)
    has_in = (indeg > 0)
    step = 1 << 22
    for lo in range(0, n, step):                                                    # bounded temporaries
        hi = min(n, lo + step)
        want = has_in[lo:hi].unsqueeze(1).to(dt) * c
        assert torch.equal(a[lo:hi], want), f"forward rows {lo}..{hi}"
    del a
    # backward: dA[u] = in_deg(u) * 2^-10 where that is exact in bf16 (in_deg < 256), else 2^-10
    gscale = 2.0 ** -10
    small = indeg < 256
    da = (torch.where(small, indeg, torch.ones_like(indeg)).to(torch.float32) * gscale).to(dt)
    da = da.unsqueeze(1).expand(n, d).contiguous()
    dq, _ = F_.edge_backward_q(g.csr, q, k, None, da, ds, ss, _lib.ACT_RELU, 0.0, False, scale_da_inplace=True)
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        # mean: dQ[u] = Σ_{deg edges} dA[u]/deg  (the kernel accumulates dA per edge and scales the row sum by 1/deg)
        want = torch.where(small[lo:hi], indeg[lo:hi], torch.ones_like(indeg[lo:hi])).to(torch.float32) * gscale
        want = (want * has_in[lo:hi]).to(dt)
        assert torch.equal(dq[lo:hi, 0], want) and torch.equal(dq[lo:hi, d - 1], want), f"dQ rows {lo}..{hi}"
    del dq
    # `da` now holds dA[u]/in_deg(u) (scaled in place): 2^-10 for the small-degree rows
    dk = F_.edge_backward_k(g.csc, q, k, None, da, None, ss, _lib.ACT_RELU, 0.0)
    # checksum of checksums in fp64: Σ_v dK[v] = Σ_u in_deg(u) · da_scaled[u]
    lhs = dk[:, 3].double().sum().item()
    rhs = (indeg.double() * da[:, 3].double()).sum().item()
    assert abs(lhs - rhs) <= 2e-3 * abs(rhs), (lhs, rhs)      # dK rows are rounded to bf16 once each
    # a sample of source rows against an fp64 gather-sum of the (scaled) dA rows they point to: one bf16 rounding
    ip = g.csc.indptr
    rows = torch.randint(0, n, (4096,), device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    rows = torch.cat([rows, outdeg.topk(8).indices])                 # ... and the largest source rows
    for v in rows.tolist()[-64:] + rows.tolist()[:192]:
        beg, end = int(ip[v]), int(ip[v + 1])
        want = da[g.csc.idx[beg:end].long(), 3].double().sum().item()
        got = float(dk[v, 3])
        assert abs(got - want) <= 2.0 ** -8 * abs(want) + 1e-30, (v, end - beg, got, want)
