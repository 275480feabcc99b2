"""Destination-row partition (sir-gcn_b200/partition.py): G ranks == 1 rank.

CPU part (`-m "not gpu"`): world_size-2 gloo run of the host logic (slicing, padding, the three all-gathers,
weight-gradient all-reduce) with a torch stand-in for the CUDA edge kernels, checked against the CPU oracle on the
whole graph.  GPU part: the same through the C-ABI kernels (world 1 on one GPU, NCCL world 2 when two are visible).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

import sirgcn_b200  # noqa: F401
from oracle.sirconv_ref import RefGraph, RefSIRConv, csr_csc_ref
from sirgcn_b200 import SIRConv, _lib, partition
from sirgcn_b200.graph import CompressedRows

ACTS = {"relu": nn.ReLU, "leaky": lambda: nn.LeakyReLU(0.2), "gelu": nn.GELU}


def _act(code, p, z):
    if code == _lib.ACT_RELU:
        return torch.relu(z), (z > 0).to(z.dtype)
    if code == _lib.ACT_LEAKY_RELU:
        return torch.where(z > 0, z, p * z), torch.where(z > 0, torch.ones_like(z), torch.full_like(z, p))
    if code == _lib.ACT_GELU:
        cdf = 0.5 * (1 + torch.erf(z * 0.7071067811865476))
        pdf = 0.3989422804014327 * torch.exp(-0.5 * z * z)
        return z * cdf, cdf + z * pdf
    return z, torch.ones_like(z)


class TorchEdgeBackend:
    """torch restatement of the three edge entry points on CompressedRows (test stand-in for the CUDA kernels)"""

    @staticmethod
    def _rid(rows):
        deg = (rows.indptr[1:] - rows.indptr[:-1]).long()
        return torch.repeat_interleave(torch.arange(rows.n_rows), deg)

    @staticmethod
    def forward(rows, q, k, e, ds, ss, act, ap, out=None, accumulate=False):
        rid, idx = TorchEdgeBackend._rid(rows), rows.idx.long()
        m, _ = _act(act, ap, q[rid] + k[idx])
        if ss is not None:
            m = m * ss[idx].to(m.dtype).unsqueeze(1)
        res = torch.zeros(rows.n_rows, q.shape[1], dtype=q.dtype).index_add_(0, rid, m)
        if ds is not None:
            res = res * ds[:rows.n_rows].to(res.dtype).unsqueeze(1)
        if out is None:
            return res
        return out.add_(res) if accumulate else out.copy_(res)

    @staticmethod
    def backward_q(rows, q, k, e, da, ds, ss, act, ap, want_de, out=None, scale_da_inplace=False):
        rid, idx = TorchEdgeBackend._rid(rows), rows.idx.long()
        _, dact = _act(act, ap, q[rid] + k[idx])
        g = da[rid] * dact
        if ds is not None:
            g = g * ds[:rows.n_rows][rid].to(g.dtype).unsqueeze(1)
        if ss is not None:
            g = g * ss[idx].to(g.dtype).unsqueeze(1)
        out.copy_(torch.zeros(rows.n_rows, q.shape[1], dtype=q.dtype).index_add_(0, rid, g))
        return out, None

    @staticmethod
    def backward_k(rows, q, k, e, da, ds, ss, act, ap, out=None):
        rid, idx = TorchEdgeBackend._rid(rows), rows.idx.long()     # rows = local sources, idx = global destinations
        _, dact = _act(act, ap, q[idx] + k[rid])
        g = da[idx] * dact
        if ds is not None:
            g = g * ds[idx].to(g.dtype).unsqueeze(1)
        if ss is not None:
            g = g * ss[:rows.n_rows][rid].to(g.dtype).unsqueeze(1)
        out.copy_(torch.zeros(rows.n_rows, k.shape[1], dtype=k.dtype).index_add_(0, rid, g))
        return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(seed, n, e, d_in, d, d_out, act, agg, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    src, dst = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    dst[: e // 3] = 2                      # a hub destination
    torch.manual_seed(seed)
    layer = RefSIRConv(d_in, d, d_out, ACTS[act](), agg_type=agg).to(dtype)
    x = torch.randn(n, d_in, dtype=dtype)
    gout = torch.randn(n, d_out, dtype=dtype)
    return src, dst, layer, x, gout


def _reference(src, dst, n, layer, x, gout):
    """layer: one RefSIRConv or a list of them applied back to back"""
    layers = layer if isinstance(layer, (list, tuple)) else [layer]
    xr = x.clone().requires_grad_(True)
    g = RefGraph(src, dst, n)
    out = xr
    for l in layers:
        out = l(g, out)
    grads = torch.autograd.grad(out, [xr] + [p for l in layers for p in l.parameters()], gout)
    return out.detach(), grads


def _gloo_worker(rank, world, port, agg, act, n_layers, chunks, use_full, gather, bwd_chunks, ret, layout=1, keep_q=None):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 37                              # not divisible by the world size: padding rows exist
        src, dst, ref, x, gout = _case(3, n, 400, 6, 12, 6, act, agg)
        torch.manual_seed(11)
        refs = [ref] + [RefSIRConv(6, 12, 6, ACTS[act](), agg_type=agg).double() for _ in range(n_layers - 1)]
        out_ref, g_ref = _reference(src, dst, n, refs, x, gout)
        c = csr_csc_ref(src, dst, n)
        csr = CompressedRows(c[0], c[1], None)
        csc = CompressedRows(c[3], c[4], None)
        part = partition.RowPartition.from_csr_csc(csr, csc, n, rank, world, layout_chunks=layout)
        assert part.transport().kind == "collective"
        assert part.n_pad % layout == 0 and part.n_pad * world >= n
        cuts = part.row_chunks(chunks)
        assert len(cuts) == chunks and cuts[0][0] == 0 and cuts[-1][1] == part.n_pad
        assert sum(r.num_pos for _, _, r in cuts if r is not None) == part.csr.num_pos
        layers = []
        for r in refs:
            layer = SIRConv(6, 12, 6, ACTS[act](), agg_type=agg).double()
            layer.load_state_dict(r.state_dict())
            layers.append(layer)
        xl = x[part.lo:part.hi].clone().requires_grad_(True)
        x_full = part.all_gather_rows(xl.detach()) if use_full else None
        if use_full:
            assert x_full.shape[0] == world * part.n_pad
            rows = torch.tensor([part.table_row(g // part.n_pad, g % part.n_pad) for g in range(n)])
            assert torch.equal(x_full[rows], x)         # gathered-table order (rank-major when layout == 1)
            if layout == 1:
                assert torch.equal(x_full[:n], x)
        out = partition.partitioned_sirconv_stack(layers, part, xl, chunks=chunks, backend=TorchEdgeBackend,
                                                  feat_full=x_full, gather=gather, bwd_chunks=bwd_chunks,
                                                  keep_q_full=keep_q)
        grads = torch.autograd.grad(out, [xl] + [p for l in layers for p in l.parameters()], gout[part.lo:part.hi])
        # degree coefficients are fp32 by design (the kernels read fp32 scales): 1e-6; pure sums: 1e-10
        tol = dict(rtol=1e-10, atol=1e-12) if agg == "sum" else dict(rtol=1e-6, atol=1e-7)
        ok = torch.allclose(out, out_ref[part.lo:part.hi], **tol)
        ok &= torch.allclose(grads[0], g_ref[0][part.lo:part.hi], **tol)
        for a, b in zip(grads[1:], g_ref[1:]):          # every rank holds the FULL-graph weight gradient
            ok &= torch.allclose(a, b, **tol)
        with torch.no_grad():
            ok &= torch.allclose(partition.partitioned_sirconv_stack(layers, part, xl, chunks=chunks, gather=gather,
                                                                     backend=TorchEdgeBackend, feat_full=x_full), out)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("agg,act,n_layers,chunks,use_full,gather,bwd_chunks", [
    ("sum", "relu", 1, 1, False, "projections", 1), ("mean", "leaky", 2, 3, False, "projections", 1),
    ("sym", "gelu", 3, 2, False, "projections", 3), ("sym", "leaky", 1, 1, True, "projections", 2),
    ("mean", "relu", 2, 4, True, "projections", 4),
    ("sym", "gelu", 3, 2, False, "inputs", 1), ("mean", "leaky", 2, 4, True, "inputs", 3),
    ("sum", "relu", 2, 1, False, "inputs", 2)])
def test_partition_gloo_world2_matches_single_process(agg, act, n_layers, chunks, use_full, gather, bwd_chunks):
    _run_gloo(2, agg, act, n_layers, chunks, use_full, gather, bwd_chunks)


@pytest.mark.parametrize("agg,act,n_layers,chunks,use_full,gather,bwd_chunks", [
    ("mean", "relu", 2, 5, True, "projections", 2), ("sym", "leaky", 2, 2, False, "inputs", 3)])
def test_partition_gloo_world3_ragged_last_rank(agg, act, n_layers, chunks, use_full, gather, bwd_chunks):
    """37 nodes over 3 ranks: n_pad = 13, the last rank owns 11 rows, and with 5 chunks some chunks of it are empty"""
    _run_gloo(3, agg, act, n_layers, chunks, use_full, gather, bwd_chunks)


def _run_gloo(world, agg, act, n_layers, chunks, use_full, gather, bwd_chunks, layout=1, keep_q=None):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, agg, act, n_layers, chunks, use_full, gather, bwd_chunks, ret, layout,
                                     keep_q), nprocs=world, join=True)
        assert dict(ret) == {r: True for r in range(world)}


@pytest.mark.parametrize("world,agg,act,n_layers,chunks,use_full,gather,bwd_chunks,layout,keep_q", [
    (2, "mean", "relu", 2, 4, True, "projections", 4, 4, None),     # cuts == layout chunks, dA under the dK walk
    (2, "mean", "relu", 2, 4, True, "projections", 1, 4, None),     # the bench's default schedule
    (2, "sym", "leaky", 2, 2, True, "projections", 1, 4, False),    # cuts of two layout chunks each
    (2, "sym", "gelu", 3, 4, False, "projections", 2, 4, None),     # no gathered input: K and Q travel
    (2, "mean", "leaky", 2, 3, True, "inputs", 3, 4, None),         # cuts that split layout chunks (partial pieces)
    (3, "sym", "relu", 2, 8, True, "projections", 8, 8, None),      # ragged last rank, empty chunks
    (3, "sum", "gelu", 2, 2, False, "inputs", 1, 2, True)])
def test_partition_gloo_chunk_major_tables(world, agg, act, n_layers, chunks, use_full, gather, bwd_chunks, layout, keep_q):
    """gathered tables laid out chunk-major (RowPartition.layout_chunks): index remap, permuted coefficient vectors,
    in-place chunk all-gathers, the one-GEMM [Q|K] projection of a gathered input kept for backward"""
    _run_gloo(world, agg, act, n_layers, chunks, use_full, gather, bwd_chunks, layout, keep_q)


@pytest.mark.parametrize("n_layers,gather,bwd_chunks,use_full", [(1, "projections", 1, False), (2, "projections", 1, True),
                                                                 (3, "inputs", 2, False)])
def test_partition_world1_host_logic(n_layers, gather, bwd_chunks, use_full):
    """a single rank (no process group): the stack function degenerates to the plain layer stack"""
    n = 41
    src, dst, ref, x, gout = _case(9, n, 500, 6, 12, 6, "leaky", "sym")
    torch.manual_seed(3)
    refs = [ref] + [RefSIRConv(6, 12, 6, ACTS["leaky"](), agg_type="sym").double() for _ in range(n_layers - 1)]
    out_ref, g_ref = _reference(src, dst, n, refs, x, gout)
    c = csr_csc_ref(src, dst, n)
    part = partition.RowPartition.from_csr_csc(CompressedRows(c[0], c[1], None), CompressedRows(c[3], c[4], None), n, 0, 1)
    layers = []
    for r in refs:
        layer = SIRConv(6, 12, 6, ACTS["leaky"](), agg_type="sym").double()
        layer.load_state_dict(r.state_dict())
        layers.append(layer)
    xl = x.clone().requires_grad_(True)
    out = partition.partitioned_sirconv_stack(layers, part, xl, chunks=3, backend=TorchEdgeBackend, gather=gather,
                                              bwd_chunks=bwd_chunks,
                                              feat_full=part.all_gather_rows(xl.detach()) if use_full else None)
    grads = torch.autograd.grad(out, [xl] + [p for l in layers for p in l.parameters()], gout)
    tol = dict(rtol=1e-6, atol=1e-7)
    assert torch.allclose(out, out_ref, **tol)
    for a, b in zip(grads, g_ref):
        assert torch.allclose(a, b, **tol)


def test_partition_bounds_cover_every_row_once():
    for n, world in [(10, 3), (37, 2), (5, 8), (16, 4)]:
        rows = []
        for r in range(world):
            n_pad, lo, hi = partition.RowPartition.bounds(n, r, world)
            assert hi - lo <= n_pad
            rows += list(range(lo, hi))
        assert rows == list(range(n))


# ---- GPU ------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


@pytest.mark.gpu
@pytest.mark.parametrize("agg,act", [("sum", "relu"), ("mean", "relu"), ("sym", "leaky")])
def test_partition_world1_equals_graph_path(agg, act):
    from sirgcn_b200 import Graph
    dev = "cuda:0"
    src, dst, ref, x, gout = _case(5, 300, 4000, 16, 64, 24, act, agg, torch.float32)
    g = Graph(src.to(dev), dst.to(dev), 300, long_threshold=64)
    layer = SIRConv(16, 64, 24, ACTS[act](), agg_type=agg).to(dev)
    layer.load_state_dict(ref.state_dict())
    xa = x.to(dev).requires_grad_(True)
    out_a = layer(g, xa)
    ga = torch.autograd.grad(out_a, [xa] + list(layer.parameters()), gout.to(dev))
    part = partition.RowPartition.from_graph(g, 0, 1)
    xb = x.to(dev).requires_grad_(True)
    out_b = partition.partitioned_sirconv(layer, part, xb)
    gb = torch.autograd.grad(out_b, [xb] + list(layer.parameters()), gout.to(dev))
    assert _rel(out_b, out_a) < 1e-5
    for a, b in zip(gb, ga):
        assert _rel(a, b) < 1e-5
    out_ref, g_ref = _reference(src, dst, 300, ref.double(), x.double(), gout.double())
    assert _rel(out_b, out_ref) < 1e-5 and _rel(gb[0], g_ref[0]) < 1e-5


@pytest.mark.gpu
def test_partition_from_hashed_generator_matches_whole_graph_slices():
    from sirgcn_b200 import Graph, synth
    dev = "cuda:0"
    n, e, world = 3001, 90000, 4
    src, dst, _ = synth.powerlaw_hashed(n, e, seed=7, device=dev)
    g = Graph(src, dst, n, long_threshold=64)
    for rank in range(world):
        whole = partition.RowPartition.from_graph(g, rank, world)
        # (synthetic_powerlaw itself needs a process group for the norm all_gather: build its pieces rank by rank)
        ip = synth.powerlaw_indptr(n, e, seed=7, device=dev)
        n_pad, lo, hi = partition.RowPartition.bounds(n, rank, world)
        in_src, in_dst = synth.powerlaw_hashed_rows(ip, n, lo, hi, seed=7)
        out_src, out_dst = synth.powerlaw_hashed_cols(ip, n, lo, hi, seed=7, step=20000)
        csr = partition.build_rows(in_dst - lo, in_src, hi - lo, 64)
        csc = partition.build_rows(out_src - lo, out_dst, hi - lo, 64)
        assert torch.equal(csr.indptr, whole.csr.indptr) and torch.equal(csr.idx, whole.csr.idx)
        assert torch.equal(csc.indptr, whole.csc.indptr) and torch.equal(csc.idx, whole.csc.idx)


def _nccl_worker(rank, world, port, transport, chunks, barrier, ret, layout=1):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["SIRGCN_TRANSPORT"], os.environ["SIRGCN_PEER_BARRIER"] = transport, barrier
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from sirgcn_b200 import synth
        n, e, d = 20011, 600000, 128
        part = partition.RowPartition.synthetic_powerlaw(n, e, rank, world, seed=3, device=dev, long_threshold=64,
                                                         layout_chunks=layout)
        assert part.transport().kind == transport
        # same device as the partition: the degree law is drawn with the device's RNG
        src, dst, _ = synth.powerlaw_hashed(n, e, seed=3, device=dev, index_dtype=torch.int64)
        src, dst = src.cpu(), dst.cpu()
        torch.manual_seed(0)
        # a smooth σ: with ReLU/LeakyReLU the fp32-vs-fp64 comparison at this size (77 M pre-activations) is dominated
        # by σ' flipping for the handful of |z| < 1e-7 elements, not by arithmetic error
        refs = [RefSIRConv(d, d, d, nn.GELU(), agg_type="mean") for _ in range(2)]
        x, gout = torch.randn(n, d), torch.randn(n, d)
        layers = []
        for r in refs:
            layer = SIRConv(d, d, d, nn.GELU(), agg_type="mean").to(dev)
            layer.load_state_dict({k: v.float() for k, v in r.state_dict().items()})
            layers.append(layer)
        out_ref, g_ref = _reference(src, dst, n, [r.double() for r in refs], x.double(), gout.double())
        params = [p for l in layers for p in l.parameters()]
        xl = x[part.lo:part.hi].to(dev).requires_grad_(True)
        for it in range(3):     # repeated steps recycle the peer slices (overwrite fences) and must stay bit-identical
            out = partition.partitioned_sirconv_stack(layers, part, xl, chunks=chunks, gather="projections")
            grads = torch.autograd.grad(out, [xl] + params, gout[part.lo:part.hi].to(dev))
            if it == 0:
                first = [out.clone()] + [g.clone() for g in grads]
            else:
                assert all(torch.equal(a, b) for a, b in zip(first, [out] + list(grads))), "step not repeatable"
        with torch.no_grad():   # inference: nothing is held for a backward that never comes
            assert torch.equal(partition.partitioned_sirconv_stack(layers, part, xl, chunks=chunks,
                                                                   gather="projections"), first[0])
        # the input gathered ahead of time (layer 1 projects its K / Q tables locally), and layer inputs instead of
        # projections travelling between the layers: same result
        x_full = part.all_gather_rows(xl.detach())
        for kw in (dict(feat_full=x_full, gather="projections"), dict(feat_full=x_full, gather="inputs"),
                   dict(gather="inputs"), dict(feat_full=x_full, bwd_chunks=3), dict(gather="inputs", bwd_chunks=2)):
            out_f = partition.partitioned_sirconv_stack(layers, part, xl, chunks=chunks, **kw)
            grads_f = torch.autograd.grad(out_f, [xl] + params, gout[part.lo:part.hi].to(dev))
            assert all(_rel(a, b) < 1e-6 for a, b in zip([out_f] + list(grads_f), first)), kw
        # layer by layer (no cross-layer prefetch, whole-table projections) agrees with the stack (the library SGEMM
        # may pick another kernel for a row chunk, so not bit for bit in fp32)
        h = xl
        for layer in layers:
            h = partition.partitioned_sirconv(layer, part, h)
        assert _rel(h, first[0]) < 1e-6
        if transport == "peer":
            part.transport().pool.check()
        errs = [_rel(out, out_ref[part.lo:part.hi]), _rel(grads[0], g_ref[0][part.lo:part.hi])]
        errs += [_rel(a, b) for a, b in zip(grads[1:], g_ref[1:])]
        ret[rank] = [float(f"{x:.3e}") for x in errs]      # out, dfeat, then (dW_Q, db_Q, dW_K, dW_R, db_R) per layer
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("transport,chunks,barrier,layout", [
    ("collective", 3, "flags", 1), ("peer", 4, "flags", 1), ("peer", 1, "nccl", 1), ("push", 4, "flags", 1),
    ("pushsm", 3, "flags", 1), ("pushtma", 4, "flags", 1),
    ("collective", 4, "flags", 4), ("push", 4, "flags", 4), ("pushsm", 2, "flags", 4), ("pushtma", 4, "flags", 4),
    ("peer", 4, "flags", 4)])
def test_partition_nccl_world2(transport, chunks, barrier, layout):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, port, transport, chunks, barrier, ret, layout), nprocs=world, join=True)
        assert len(ret) == 2 and max(max(v) for v in ret.values()) < 1e-5, dict(ret)


# ---- batched small graphs: plain data parallelism -------------------------------------
def _ddp_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from torch.nn.parallel import DistributedDataParallel as DDP
        from sirgcn_b200 import Graph, SIREConv, synth

        def model():
            torch.manual_seed(0)
            return nn.ModuleList([SIREConv(5 if i == 0 else 64, 1, 64, 64, nn.LeakyReLU(0.2), agg_type="sum")
                                  for i in range(2)]).to(dev)

        def batch(seed):
            src, dst, n, pos, dist_e = synth.cifar_like(num_graphs=16, seed=seed)
            g = torch.Generator().manual_seed(seed)
            x = torch.cat([torch.rand(n, 3, generator=g), pos], 1)
            return Graph(src.to(dev), dst.to(dev), n), x.to(dev), dist_e.unsqueeze(1).to(dev)

        def loss_of(m, b):
            g, h, ef = b
            for conv in m:
                h = conv(g, h, ef)
            return h.square().mean()

        class Wrap(nn.Module):          # DDP wants one module whose forward returns the loss inputs
            def __init__(self, convs):
                super().__init__()
                self.convs = convs

            def forward(self, b):
                return loss_of(self.convs, b)

        ddp = DDP(Wrap(model()), device_ids=[rank])
        ddp(batch(100 + rank)).backward()                       # this rank's own 16-graph batch
        got = [p.grad.clone() for p in ddp.parameters()]
        ref = Wrap(model())                                     # one process, both batches, averaged
        sum(loss_of(ref.convs, batch(100 + r)) for r in range(world)).div(world).backward()
        errs = [_rel(a, p.grad) for a, p in zip(got, ref.parameters())]
        ret[rank] = max(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_data_parallel_batches_world2():
    """config C run data-parallel: per-rank batches of kNN graphs, replicated weights, DDP's gradient all-reduce —
    the layers are ordinary autograd Functions, so DistributedDataParallel applies unchanged"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == 2 and max(ret.values()) < 1e-5, dict(ret)
