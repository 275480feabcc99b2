#!/usr/bin/env python
"""bench.py — SIR-GCN conv fwd+bwd throughput (Gedges/s) and fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload P|A|Z|C] [--dtype f32|bf16] [--scale S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on host cores

One "step" = forward + backward of the L SIRConv layers of the workload over the whole synthetic
graph (BASELINE.json configs[4] "P": power-law graph, 50 M nodes / 2 B edges, d_hidden 128, 2 layers,
bf16 tables with fp32 accumulation, σ = ReLU, mean aggregation; destination-row partitioned when
N > 1).  Metric: Gedges/s = E·L / t(step).  `value` is timed with the node features resident in HBM;
`e2e` times the same step through the layer API with the node features copied from pinned host memory
and a checksum of the output + the weight gradients read back, every step.
Inputs (≥ 12.8 GB per table at scale 1) are far larger than the 126 MB L2, so no L2 flush is needed.

The `roofline` object is measured live: CUDA events are recorded around every sirgcn_edge_* C-ABI call
(on the stream the kernels are launched on) inside the timed region; algorithmic bytes follow
SURVEY.md §8(d) / DESIGN.md.  `cpu_baseline` is the CPU oracle (a port of the reference's DGL path,
oracle/sirconv_ref.py) timed on a bounded sample of the same workload on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
# N > 1: compute, barrier, NCCL and 7 copy streams — more hardware queues than the default 8 avoids false ordering
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402
from torch import nn  # noqa: E402

_REAL_STDOUT = None
METRIC = "sirgcn_conv_fwd_bwd_gedges_per_s"
UNIT = "Gedges/s"

WORKLOADS = {
    # name: nodes, edges, d_in, d_hidden, layers, dtype, agg, act
    "P": dict(nodes=50_000_000, edges=2_000_000_000, d_in=128, d=128, layers=2, dtype="bf16", agg="mean",
              act="relu", max_deg=None, desc="power-law 50M nodes / 2B edges, d=128, 2 layers (BASELINE configs[4])"),
    "A": dict(nodes=169_343, edges=1_166_243, d_in=128, d=256, layers=3, dtype="f32", agg="sum",
              act="leaky", max_deg=13_000, desc="ogbn-arxiv-shaped 169,343 nodes / 1.17M edges, 128->256, 3 layers (configs[2])"),
    # batched small graphs: a NEW batch (graph conversion included) every step, data-parallel over the GPUs
    "Z": dict(nodes=128 * 23, edges=128 * 50, d_in=64, d=64, layers=4, dtype="f32", agg="sum", act="leaky",
              max_deg=None, edge_types=4,
              desc="ZINC-shaped batch: 128 molecular graphs x 23 nodes / 50 edges, bond-type edge term, d=64, 4 layers (configs[1])"),
    "C": dict(nodes=None, edges=None, d_in=5, d=128, layers=4, dtype="f32", agg="sum", act="leaky", max_deg=None,
              desc="CIFAR10-super-pixel-shaped batch: 128 kNN graphs (k=8, 85..150 nodes), 5->128, 4 layers (configs[3])"),
}
BATCHED = ("Z", "C")
DTYPES = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}


def make_act(name):
    return {"relu": nn.ReLU(), "leaky": nn.LeakyReLU(0.2), "gelu": nn.GELU()}[name]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def edge_bytes(kind, E, N, r, eps=0):
    """algorithmic bytes of one edge-stage launch: r = d·sizeof(elem)"""
    if kind == "sirgcn_edge_fwd":
        return E * (r + 4 + eps) + N * (2 * r + 4)
    if kind == "sirgcn_edge_bwd_q":
        return E * (r + 4 + eps) + N * (3 * r + 4)
    return E * (2 * r + 4 + eps) + N * (2 * r + 4)


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle port of the reference's DGL path, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample_shape(w, target_edges):
    f = max(1, round(w["edges"] / target_edges))
    return max(64, w["nodes"] // f), max(64, w["edges"] // f), f


def small_batch(name, w, seed):
    """one synthetic input of a small workload on the host: dict(src, dst int32, n, x fp32 [n, d_in], etype int64 | None)"""
    from sirgcn_b200 import synth
    g = torch.Generator().manual_seed(1000 + seed)
    if name == "Z":
        src, dst, n, _, bond = synth.zinc_like(num_graphs=128, seed=seed)
        return dict(src=src.int(), dst=dst.int(), n=n, x=torch.randn(n, w["d_in"], generator=g), etype=bond)
    if name == "C":
        src, dst, n, pos, _ = synth.cifar_like(num_graphs=128, seed=seed)
        return dict(src=src.int(), dst=dst.int(), n=n, x=torch.cat([torch.rand(n, 3, generator=g), pos], 1), etype=None)
    src, dst, n = synth.arxiv_like(seed=seed)
    return dict(src=src.int(), dst=dst.int(), n=n, x=torch.randn(n, w["d_in"], generator=g), etype=None)


def small_layers(w, sir, sire, edge_types=None):
    """the L conv layers of a small workload from the given classes (CUDA package or oracle)"""
    dims = [w["d_in"]] + [w["d"]] * w["layers"]
    layers = []
    for i in range(w["layers"]):
        if edge_types:
            c = sire(dims[i], edge_types, w["d"], w["d"], make_act(w["act"]), agg_type=w["agg"])
            c.linear_edge = nn.Embedding(edge_types, w["d"])        # benchmark-datasets/zinc/model.py:12-15
        else:
            c = sir(dims[i], w["d"], w["d"], make_act(w["act"]), agg_type=w["agg"])
        layers.append(c)
    return nn.ModuleList(layers)


def build_cpu_case(args, w):
    """(step function, edges x layers per step, description of the sample) of the CPU arm: the oracle port"""
    from oracle.sirconv_ref import RefGraph, RefSIRConv, RefSIREConv
    from sirgcn_b200 import synth
    torch.manual_seed(0)
    if args.workload == "P":
        n, e, f = cpu_sample_shape(w, args.cpu_edges)
        src, dst, n = synth.powerlaw(n, e, alpha=2.3, max_deg=w["max_deg"], seed=0, device="cpu", index_dtype=torch.int64)
        x, etype = torch.randn(n, w["d_in"]), None
        sample = (f"1/{f} sample of the workload with the same degree law: {n:,} nodes / {e:,} edges, fp32, "
                  f"{w['layers']} layers fwd+bwd, oracle/sirconv_ref.py (index_select / index_add_, the ops DGL lowers to)")
        shape = {"nodes": n, "edges": e, "fraction": f"1/{f}"}
    else:
        b = small_batch(args.workload, w, 0)
        src, dst, n, x, etype = b["src"].long(), b["dst"].long(), b["n"], b["x"], b["etype"]
        e = int(src.numel())
        sample = (f"the whole workload: {n:,} nodes / {e:,} edges, fp32, {w['layers']} layers fwd+bwd, "
                  f"oracle/sirconv_ref.py (index_select / index_add_, the ops DGL lowers to)")
        shape = {"nodes": n, "edges": e, "fraction": "1/1"}
    layers = small_layers(w, RefSIRConv, RefSIREConv, w.get("edge_types"))
    gout = torch.randn(n, w["d"])
    g = RefGraph(src, dst, n)

    def step():
        h = x.clone().requires_grad_(True)
        for p in layers.parameters():
            p.grad = None
        out = h
        for l in layers:
            out = l(g, out, etype) if etype is not None else l(g, out)
        out.backward(gout)
        return float(out.detach().sum())

    return step, e * w["layers"], sample, shape


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, edges_per_step, sample, shape = build_cpu_case(args, w)
    t = time_cpu(step, args.steps, args.warmup)
    value = edges_per_step / t / 1e9
    cfg = workload_config(args, w)
    # the CPU arm times a BOUNDED SAMPLE of the workload named above: say so in the config itself
    cfg["reference_sample"] = dict(shape, dtype="f32", note="what this line actually timed; `nodes`/`edges` above name "
                                   "the workload the sample is drawn from")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.workload in BATCHED else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_config(args, w):
    if args.workload != "P":
        c_nodes = w["nodes"] if w["nodes"] else "128 graphs x U{85..150} (~15.0 k per batch)"
        c_edges = w["edges"] if w["edges"] else "8 x nodes (~120 k per batch)"
        per_step = ("a NEW batch every step: COO -> CSR/CSC conversion (Graph) and DropEdge(0) per layer are inside the "
                    "timed region" if args.workload in BATCHED else "static graph, converted once")
        return {"workload": f"{args.workload}: {w['desc']}", "nodes": c_nodes, "edges": c_edges, "d_in": w["d_in"],
                "d_hidden": w["d"], "layers": w["layers"], "agg": w["agg"], "activation": w["act"],
                "table_dtype": w["dtype"], "per_step": per_step,
                "partition": "single GPU" if args.gpus == 1 else
                (f"data parallel over {args.gpus} GPUs: one batch per GPU per step, replicated weights, ONE flat NCCL "
                 f"all-reduce of the weight gradients per step" if args.workload in BATCHED else
                 f"{args.gpus} independent replicas (the graph fits one GPU: replicas only, no collective)"),
                "l2": "L2 flushed (256 MiB write) between timed steps"}
    n, e = scaled_shape(args, w)
    return {"workload": f"{args.workload}: {w['desc']}" + ("" if args.scale == 1 else f" at scale {args.scale:g}"),
            "nodes": n, "edges": e, "d_in": w["d_in"], "d_hidden": w["d"], "layers": w["layers"],
            "agg": w["agg"], "activation": w["act"], "table_dtype": w["dtype"],
            "partition": "single GPU" if args.gpus == 1 else
            f"1-D destination rows over {args.gpus} GPUs; row tables travel by {transport_name(args)}; "
            + (f"layer l+1's INPUT rows gathered in {args.chunks} destination chunks under layer l's walk and its K/Q "
               f"tables projected locally by every rank, dA gathered under the dQ walk"
               if args.gather == "inputs" else
               f"layer l+1's K gathered in {args.chunks} destination chunks under layer l's walk, Q and dA gathers "
               f"under the other walks")
            + ("" if args.bwd_chunks <= 1 else f"; dK walk in {args.bwd_chunks} source chunks with the next dA table travelling under it")
            + (f"; gathered tables chunk-major ({layout_chunks(args)} chunks: a chunk of all ranks lands in place, no staging copy)"
               if layout_chunks(args) > 1 else "")
            + ("" if args.no_input_gather else "; node features of all ranks gathered ahead of the layers (layer 1 projects K/Q locally)"),
            "l2": "inputs >> 126 MB L2, no flush" if e * w["d"] * (4 if w["dtype"] == "f32" else 2) > (1 << 30)
            else "L2 flushed (256 MiB write) between timed steps"}


def layout_chunks(args):
    """gathered tables are laid out chunk-major with as many chunks as the cross-layer prefetch uses (partition.py)"""
    return args.layout_chunks if args.layout_chunks > 0 else max(args.chunks, args.bwd_chunks, 1)


def transport_name(args):
    kind = os.environ.get("SIRGCN_TRANSPORT", args.transport)
    if kind == "auto":
        from sirgcn_b200 import partition
        kind = partition.auto_transport(args.gpus)
    return {"peer": "copy-engine pulls from IPC-mapped peer slices",
            "push": "copy-engine pushes into IPC-mapped peer tables",
            "pushsm": "a fan-out push kernel writing into IPC-mapped peer tables",
            "pushtma": "a TMA bulk-copy fan-out kernel writing into IPC-mapped peer tables",
            "collective": "NCCL all-gathers"}[kind]


def scaled_shape(args, w):
    return max(64, int(w["nodes"] * args.scale)), max(64, int(w["edges"] * args.scale))


# ------------------------------------------------------------------------------------------------
# parity gate of the partitioned path (N > 1), run before anything is timed
# ------------------------------------------------------------------------------------------------
def parity_verdict(gated, loose, tol, tol_loose, device):
    """(worst gated error over ALL ranks, passed): every rank contributes what it measured — rank 0 alone holds the
    oracle leg — and every rank gets the SAME answer, so that all of them take the same branch afterwards (a rank that
    exits alone leaves the others hanging in the next collective: round 2 lost its 8-GPU budget to exactly that)."""
    import torch.distributed as dist
    ratio = max([v / tol for v in gated.values()] + [v / tol_loose for v in loose.values()])
    t = torch.tensor([max(gated.values()), ratio], device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), bool(float(t[1]) <= 1.0)


def partition_parity_check(args, w, layers, dev, rank, world):
    """The L-layer stack on a ~1 M-edge hashed power-law graph through THE transport and schedule that will be timed,
    against (a) the single-rank kernels on the whole graph (every rank) and (b) the fp64 CPU oracle (rank 0): output
    rows, input-gradient rows and every weight gradient.  Returns the dict that goes into the JSON line; the caller
    exits non-zero above 2e-2 (the bf16 tolerance of north_star)."""
    import torch.distributed as dist
    from sirgcn_b200 import Graph, partition, synth
    dtype = DTYPES[w["dtype"]]
    n, e, d = 100_000, 1_000_000, w["d"]
    src, dst, _ = synth.powerlaw_hashed(n, e, alpha=2.3, max_deg=None, seed=3, device=dev)
    whole = Graph(src, dst, n, need_eid=False)
    part = partition.RowPartition.synthetic_powerlaw(n, e, rank, world, alpha=2.3, max_deg=None, seed=3, device=dev,
                                                     transport=args.transport, layout_chunks=layout_chunks(args))
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)                                   # the same features on every rank
    x = torch.randn((n, w["d_in"]), generator=gen, device=dev).to(dtype)
    gout = torch.randn((n, d), generator=gen, device=dev).to(dtype)
    params = list(layers.parameters())

    def grads_of(fn, xin, g):
        for p in params:
            p.grad = None
        xin = xin.detach().clone().requires_grad_(True)
        out = fn(xin)
        out.backward(g)
        return out.detach(), xin.grad, [p.grad.detach().clone() for p in params]

    def single(h):
        for layer in layers:
            h = layer(whole, h)
        return h

    out1, dx1, dw1 = grads_of(single, x, gout)
    lo, hi = part.lo, part.hi
    feat_full = None if args.no_input_gather else part.all_gather_rows(x[lo:hi])
    outp, dxp, dwp = grads_of(lambda h: partition.partitioned_sirconv_stack(
        list(layers), part, h, chunks=args.chunks, gather=args.gather, bwd_chunks=args.bwd_chunks,
        feat_full=feat_full), x[lo:hi], gout[lo:hi])

    def rel(a, b):          # max-abs error relative to the tensor max (the metric of the parity tests)
        a, b = a.double(), b.double()
        return float((a - b).abs().max() / b.abs().max().clamp(min=1e-20)) if b.numel() else 0.0

    def fro(a, b):          # Frobenius-relative error
        a, b = a.double(), b.double()
        return float((a - b).norm() / b.norm().clamp(min=1e-20)) if b.numel() else 0.0

    # (a) the gate proper: G ranks == 1 rank through the SAME kernels (same rounding points), max-abs, every rank
    errs = {"out": rel(outp, out1[lo:hi]), "dX": rel(dxp, dx1[lo:hi]),
            "dW": max(rel(a, b) for a, b in zip(dwp, dw1))}
    res = {"graph": f"hashed power-law {n:,} nodes / {e:,} edges, {w['layers']} layers, {w['dtype']}",
           "transport": part.transport().kind, "world": world, "vs_single_rank_kernels": errs}
    gated, loose = dict(errs), {}
    if rank == 0:
        # (b) against the fp64 CPU oracle (rank 0's rows).  The forward output is gated max-abs.  Gradients of a
        # 16-bit run through a σ' that jumps at 0 (ReLU) are reported in both metrics but gated in the Frobenius
        # norm only: a 16-bit run rounds its stored tables, a handful of pre-activations land on the other side of
        # 0 than in fp64, and each such flip moves single gradient elements by a few % of the tensor max (measured:
        # 0.06 max-abs on dX at 8 GPUs with the partitioned result equal to the single-rank kernels' to 4.5e-3).
        from oracle.sirconv_ref import RefGraph, RefSIRConv
        torch.set_num_threads(os.cpu_count() or 1)
        ref = small_layers(w, RefSIRConv, None).double()
        # what the reference run in the table dtype sees: weights and stored tables rounded to it (oracle header)
        ref.load_state_dict({k: (v.detach().to(dtype) if v.dim() > 1 else v.detach()).cpu().double()
                             for k, v in layers.state_dict().items()})     # matrices are cast per call, biases stay fp32
        for l in ref:
            l.storage_dtype = None if dtype == torch.float32 else dtype
        xr = x.detach().cpu().double().requires_grad_(True)
        h = xr
        rg = RefGraph(src.cpu().long(), dst.cpu().long(), n)
        for l in ref:
            h = l(rg, h)
        gr = torch.autograd.grad(h, [xr] + list(ref.parameters()), gout.cpu().double())
        res["vs_fp64_oracle"] = {
            "out": rel(outp.cpu(), h.detach()[lo:hi]),
            "dX": rel(dxp.cpu(), gr[0][lo:hi]), "dW": max(rel(a.cpu(), b) for a, b in zip(dwp, gr[1:])),
            "dX_frobenius": fro(dxp.cpu(), gr[0][lo:hi]), "dW_frobenius": max(fro(a.cpu(), b) for a, b in zip(dwp, gr[1:]))}
        gated["oracle_out"] = res["vs_fp64_oracle"]["out"]
        if dtype == torch.float32:
            gated["oracle_dX"], gated["oracle_dW"] = res["vs_fp64_oracle"]["dX"], res["vs_fp64_oracle"]["dW"]
        else:
            loose["oracle_dX_frobenius"] = res["vs_fp64_oracle"]["dX_frobenius"]
            loose["oracle_dW_frobenius"] = res["vs_fp64_oracle"]["dW_frobenius"]
        del ref, xr, h, gr, rg
    # ONE verdict for all ranks (every rank must take the same branch afterwards): the worst error / tolerance ratio
    tol = 2e-2 if dtype != torch.float32 else 1e-5
    tol_loose = 5e-2        # 16-bit gradients against fp64 in the Frobenius norm: catches a wrong result, not rounding
    res["max_rel_err"], res["passed"] = parity_verdict(gated, loose, tol, tol_loose, dev)
    res["gated"] = sorted(gated)
    res["tolerance"] = tol
    res["gated_frobenius"], res["tolerance_frobenius"] = sorted(loose), tol_loose
    for p in params:
        p.grad = None
    del whole, part, x, gout, out1, dx1, dw1, outp, dxp, dwp
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
# small workloads: Z / C (a new batch of small graphs every step, data parallel) and A (one static graph)
# ------------------------------------------------------------------------------------------------
def run_small(args, w):
    import torch.distributed as dist
    import sirgcn_b200  # noqa: F401  (fails loudly if libsirgcn.so is missing: no CPU fallback)
    from sirgcn_b200 import DropEdge, Graph, SIRConv, SIREConv, _lib, function

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); use --impl reference")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    name, batched = args.workload, args.workload in BATCHED
    dp = batched and world > 1
    dtype = DTYPES[w["dtype"]]
    L, d = w["layers"], w["d"]
    torch.manual_seed(0)
    layers = small_layers(w, SIRConv, SIREConv, w.get("edge_types")).to(dev)
    if dtype != torch.float32:
        layers = layers.to(dtype)
    if world > 1:
        for p in layers.parameters():
            dist.broadcast(p.data, 0)
    params = list(layers.parameters())
    drop = DropEdge(0.0)          # the reference calls DropEdge per layer per step even at p = 0 (zinc/model.py:50)

    # a pool of host batches (pinned): a different one per step and per rank for Z / C, the one graph for A
    pool = [small_batch(name, w, 17 * rank + i) for i in range(4 if batched else 1)]
    for b in pool:
        b["x"] = b["x"].to(dtype)
        b["gout"] = torch.randn(b["n"], d, generator=torch.Generator().manual_seed(5)).to(dtype)
        for k in ("src", "dst", "x", "gout", "etype"):
            if b[k] is not None:
                b[k] = b[k].pin_memory()
    edges_per_step = sum(int(b["src"].numel()) for b in pool) / len(pool)
    nodes_per_step = sum(b["n"] for b in pool) / len(pool)

    def to_dev(b):
        return {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in b.items()}

    resident = [to_dev(b) for b in pool]
    need_eid = bool(w.get("edge_types")) or w["agg"] in ("max", "min")      # edge ids: edge features and the split path
    static_graph = None if batched else Graph(resident[0]["src"], resident[0]["dst"], resident[0]["n"], need_eid=need_eid)

    def step(b):
        for p in params:
            p.grad = None
        g = static_graph if static_graph is not None else Graph(b["src"], b["dst"], b["n"], need_eid=need_eid)
        h = b["x"]
        for layer in layers:
            gl, ef = drop(g, b["etype"])
            h = layer(gl, h, ef) if ef is not None else layer(gl, h)
        check = h.detach().sum(dtype=torch.float32)
        h.backward(b["gout"])
        if dp:      # ONE flat all-reduce of every weight gradient
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                p.grad.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        return check

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(1 << 28, dtype=torch.uint8, device=dev)      # 256 MiB > 126 MB L2

    def timed(fn, steps):
        barrier()
        evs = []
        for i in range(steps):
            flush_buf.fill_(1)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            fn(i)
            t1.record()
            evs.append((t0, t1))
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(args.warmup):
        step(resident[i % len(resident)])
    sampler = ClockSampler(local) if rank == 0 else None
    function.EDGE_TIMERS = []
    launches0 = _lib.launch_count()
    ms_step = timed(lambda i: step(resident[i % len(resident)]), args.steps)
    launches = _lib.launch_count() - launches0
    timers, function.EDGE_TIMERS = function.EDGE_TIMERS, None
    clocks = sampler.stop() if sampler else None

    es = torch.empty((), dtype=dtype).element_size()
    per = {}
    for fn_name, t0, t1, rows in timers:
        acc = per.setdefault(fn_name, [0.0, 0, 0])
        acc[0] += t0.elapsed_time(t1)
        acc[1] += 1
        acc[2] += edge_bytes(fn_name, rows[0], rows[1], d * es, d * es if w.get("edge_types") else 0)
    peak, peak_src = peaks()
    stages = {}
    for fn_name, (ms, cnt, by_all) in per.items():
        by, avg = by_all / cnt, ms / cnt
        stages[fn_name] = {"ms": avg, "bytes": by, "calls_per_step": cnt / args.steps, "gbs": by / avg / 1e6,
                           "frac": by / avg / 1e6 / peak, "share_of_step": ms / (ms_step * args.steps)}
    dom = max(stages, key=lambda k: per[k][0]) if stages else None

    # ---- e2e: the batch (COO, features, bond types) from pinned host memory, loss + weight gradients read back
    d2h = [0]

    def e2e_step(i):
        b = to_dev(pool[i % len(pool)])
        check = step(b)
        outs = [check.cpu()] + [p.grad.cpu() for p in params]
        d2h[0] = sum(t.numel() * t.element_size() for t in outs)

    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    h2d = sum(sum(v.numel() * v.element_size() for v in b.values() if torch.is_tensor(v)) for b in pool) / len(pool)
    if not batched:
        h2d -= (pool[0]["src"].numel() + pool[0]["dst"].numel()) * 4     # the static graph is converted once

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cstep, cedges, sample, _ = build_cpu_case(args, w)
        t = time_cpu(cstep, 2, 1)
        cpu = {"value": cedges / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        total_edges = edges_per_step * L * world
        line = {
            "metric": METRIC, "value": total_edges / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "us_per_step": ms_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": dict(workload_config(args, w), nodes_per_step=nodes_per_step, edges_per_step=edges_per_step),
            "e2e": {"value": total_edges / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * world,
                    "d2h_bytes_per_step": d2h[0] * world, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": None if dom is None else {
                "bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": None, "peak_source": peak_src,
                "note": "working set fits the 126 MB L2: latency / launch-bound, read ms_per_step (SURVEY.md 8d)",
                "stages": stages},
            "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args, w):
    import torch.distributed as dist
    import sirgcn_b200  # noqa: F401  (fails loudly if libsirgcn.so is missing: no CPU fallback)
    from sirgcn_b200 import Graph, SIRConv, _lib, function, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    dtype = DTYPES[w["dtype"]]
    n, e = scaled_shape(args, w)
    L, d = w["layers"], w["d"]
    torch.manual_seed(0)
    dims = [w["d_in"]] + [d] * L
    layers = nn.ModuleList([SIRConv(dims[i], d, d, make_act(w["act"]), agg_type=w["agg"]) for i in range(L)]).to(dev)
    if world > 1:
        for p in layers.parameters():
            dist.broadcast(p.data, 0)
    params = list(layers.parameters())
    if dtype != torch.float32:
        pass        # weights stay fp32 (master copies); the layers cast them to the table dtype per call

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = partition_parity_check(args, w, layers, dev, rank, world)
        if not parity["passed"]:                # the same verdict on every rank (all-reduced inside)
            if rank == 0:
                sys.stderr.write(f"parity check FAILED before timing: {json.dumps(parity)}\n")
            dist.destroy_process_group()
            raise SystemExit(3)

    # ---- graph + features -------------------------------------------------------------------
    t_build = time.perf_counter()
    if world == 1:
        # counter-based generator: the SAME global graph whatever the world size (partition.py slices it)
        src, dst, n = synth.powerlaw_hashed(n, e, alpha=2.3, max_deg=w["max_deg"], seed=0, device=dev)
        graph = Graph(src, dst, n, need_eid=False, keep_coo=False)
        del src, dst
        n_local, e_local = n, e

        def run_layers(h):
            for layer in layers:
                h = layer(graph, h)
            return h
    else:
        from sirgcn_b200 import partition
        part = partition.RowPartition.synthetic_powerlaw(n, e, rank, world, alpha=2.3, max_deg=w["max_deg"],
                                                         seed=0, device=dev, transport=args.transport,
                                                         layout_chunks=layout_chunks(args))
        n_local, e_local = part.n_local, part.num_local_edges
        # the node features of ALL ranks, gathered ahead of the layers (an input: a prefetching loader gathers step
        # i+1's under step i); layer 1 then projects its K / Q tables locally instead of gathering both in line
        full_of = {}

        def run_layers(h):
            return partition.partitioned_sirconv_stack(list(layers), part, h, chunks=args.chunks, gather=args.gather,
                                                       bwd_chunks=args.bwd_chunks,
                                                       feat_full=full_of.get(h.data_ptr()))
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    x_dev = torch.empty((n_local, w["d_in"]), dtype=dtype, device=dev)
    chunk = 1 << 22
    for lo in range(0, n_local, chunk):   # bounded fp32 temporaries
        hi = min(n_local, lo + chunk)
        x_dev[lo:hi] = torch.randn((hi - lo, w["d_in"]), generator=gen, device=dev).to(dtype)
    gout = torch.empty((n_local, d), dtype=dtype, device=dev)
    for lo in range(0, n_local, chunk):
        hi = min(n_local, lo + chunk)
        gout[lo:hi] = torch.randn((hi - lo, d), generator=gen, device=dev).to(dtype)
    x_dev.requires_grad_(True)
    if world > 1 and not args.no_input_gather:
        full_of[x_dev.data_ptr()] = part.all_gather_rows(x_dev.detach())

    def step(x):
        x.grad = None
        for p in params:
            p.grad = None
        h = run_layers(x)
        check = h.detach().sum(dtype=torch.float32)
        h.backward(gout)
        return check

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    small = e * d * torch.empty((), dtype=dtype).element_size() <= (1 << 30)
    flush_buf = torch.empty(1 << 28, dtype=torch.uint8, device=dev) if small else None   # 256 MiB > 126 MB L2

    def timed(fn, steps):
        """K steps between barriers; each step bracketed by CUDA events on the launching stream (the L2
        flush of small workloads sits between the brackets); returns the max over ranks of ms/step"""
        barrier()
        evs = []
        for _ in range(steps):
            if flush_buf is not None:
                flush_buf.fill_(1)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            fn()
            t1.record()
            evs.append((t0, t1))
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    for _ in range(args.warmup):
        step(x_dev)
    sampler = ClockSampler(local) if rank == 0 else None
    function.EDGE_TIMERS = []
    if world > 1:
        partition.PHASE_MARKS = []
    launches0 = _lib.launch_count()
    # under `ncu --profile-from-start off` only the timed region is captured (graph generation and conversion launch
    # hundreds of 2 B-element kernels whose save/restore makes a whole-program capture take half an hour)
    torch.cuda.profiler.start()
    ms_step = timed(lambda: step(x_dev), args.steps)
    torch.cuda.profiler.stop()
    launches = _lib.launch_count() - launches0
    timers, function.EDGE_TIMERS = function.EDGE_TIMERS, None
    phases = None
    if world > 1:
        marks, partition.PHASE_MARKS = partition.PHASE_MARKS, None
        acc = {}
        for (l0, e0), (l1, e1) in zip(marks[:-1], marks[1:]):
            if l1.endswith(":start"):
                continue
            acc[l1] = acc.get(l1, 0.0) + e0.elapsed_time(e1)
        phases = {k: v / args.steps for k, v in acc.items()}      # ms per step (all layers), rank 0
    clocks = sampler.stop() if sampler else None
    peak_mem = torch.cuda.max_memory_allocated() / 2**30
    if os.environ.get("SIRGCN_BENCH_VALUE_ONLY"):       # profiling / A-B runs: stop after the timed region
        if rank == 0:
            emit({"metric": METRIC, "value": e * L / (ms_step * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_step,
                  "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "gpu_launches": launches,
                  "config": workload_config(args, w), "phases_ms_per_step": phases, "parity_check": parity,
                  "clocks": clocks, "peak_mem_gib": peak_mem,
                  "note": "SIRGCN_BENCH_VALUE_ONLY: value leg only (A/B or profiling run), not a bench line"})
        if world > 1:
            dist.destroy_process_group()
        return

    # per-kernel durations (CUDA events around each C-ABI edge call, inside the timed region)
    es = torch.empty((), dtype=dtype).element_size()
    per = {}
    for name, t0, t1, rows in timers:
        acc = per.setdefault(name, [0.0, 0, 0])
        acc[0] += t0.elapsed_time(t1)
        acc[1] += 1
        acc[2] += edge_bytes(name, rows[0], rows[1], d * es)
    peak, peak_src = peaks()
    stages = {}
    for name, (ms, cnt, by_all) in per.items():
        by, avg = by_all / cnt, ms / cnt            # per C-ABI call (a chunked walk is several calls)
        stages[name] = {"ms": avg, "bytes": by, "calls_per_step": cnt / args.steps, "gbs": by / avg / 1e6,
                        "frac": by / avg / 1e6 / peak, "share_of_step": ms / (ms_step * args.steps)}
    dom = max(stages, key=lambda k: per[k][0]) if stages else None
    traffic = None      # dram bytes per launch of the dominant kernel from the committed ncu --set full capture
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    tpath = tfiles[-1] if tfiles else ""
    if dom and world == 1 and tpath:
        with open(tpath) as f:
            tj = json.load(f)
        if tj["workload"] == {"nodes": n, "edges": e, "d": d, "dtype": w["dtype"]}:
            traffic = tj.get(dom)
    tot_ms = sum(v[0] for v in per.values()) / (args.steps * L)        # per layer: one forward + two backward walks
    tot_by = sum(v[2] for v in per.values()) / (args.steps * L)

    # ---- e2e: node features from pinned host memory, results read back, every step -----------------
    x_dev.grad = None
    x_host = torch.empty((n_local, w["d_in"]), dtype=dtype, pin_memory=True)
    x_host.copy_(x_dev.detach())
    d2h = [0]

    # Double-buffered input pipeline when HBM allows it: step i+1's features travel on a copy stream while step i
    # computes (what a training loop with a prefetching loader does); every step's copy is inside the timed region.
    nbytes_x = x_host.numel() * x_host.element_size()
    headroom = torch.cuda.mem_get_info(dev)[1] - torch.cuda.max_memory_reserved(dev)
    # host-resident inputs: gathering them over NVLink every step as well (11.2 GB per rank at 8 GPUs, on the same
    # NCCL queue as the layers' own table traffic) costs more than it saves — measured 230 vs 185-200 ms per step at
    # 8 GPUs — so by default the e2e leg lets layer 1 gather its K and Q tables in line instead
    gather_input = world > 1 and not args.no_input_gather and args.e2e_input_gather
    if world > 1 and not gather_input:
        full_of.clear()
    extra = nbytes_x * (1 + (world if gather_input else 0))     # a second input buffer (+ its gathered copy)
    double = headroom > int(1.25 * extra) + (2 << 30)
    bufs = [x_dev, torch.empty_like(x_dev).requires_grad_(True)] if double else [x_dev]
    if gather_input:
        for b in bufs[1:]:
            full_of[b.data_ptr()] = torch.empty_like(full_of[x_dev.data_ptr()])
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in bufs]        # H2D of buffer b finished
    freed = [torch.cuda.Event() for _ in bufs]        # last step that read buffer b finished
    state = {"i": 0, "primed": False}

    def prefetch(b):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(freed[b])
            bufs[b].copy_(x_host, non_blocking=True)   # this step's inputs: pinned host -> HBM
            if gather_input:                            # ... and over NVLink to every rank, still on the copy stream
                part.all_gather_rows(bufs[b].detach(), out=full_of[bufs[b].data_ptr()])
            ready[b].record(copy_stream)

    def e2e_step():
        b = state["i"] % len(bufs)
        if not state["primed"] or not double:
            freed[b].record()
            prefetch(b)
            state["primed"] = True
        if double:
            nb = (state["i"] + 1) % len(bufs)
            freed[nb].record()                          # everything enqueued so far (incl. step i-1 on nb) precedes
            prefetch(nb)                                # the next step's copy, which overlaps this step
        torch.cuda.current_stream(dev).wait_event(ready[b])
        check = step(bufs[b])
        outs = [check.cpu()] + [p.grad.cpu() for p in params]   # results back on the host
        d2h[0] = sum(t.numel() * t.element_size() for t in outs)
        bufs[b].grad = None
        state["i"] += 1

    e2e_step()
    ms_e2e = timed(e2e_step, max(1, min(args.steps, 3)))
    h2d = x_host.numel() * x_host.element_size()

    if world > 1:
        pool = getattr(part.transport(), "pool", None)
        if pool is not None:
            pool.check()            # a device-side barrier or bulk copy that gave up must not pass silently: no line

    # ---- CPU baseline on rank 0 (bounded sample) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del x_host
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cstep, cedges, sample, _ = build_cpu_case(args, w)
        t = time_cpu(cstep, 2, 1)
        cpu = {"value": cedges / t / 1e9, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        total_edges = e * L
        line = {
            "metric": METRIC, "value": total_edges / (ms_step * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": workload_config(args, w),
            "e2e": {"value": total_edges / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h[0] * world,
                    "ms_per_step": ms_e2e, "input_pipeline": ("double-buffered H2D on a copy stream" if double
                    else "single buffer, H2D on the compute stream") + ("" if world == 1 else
                    ("; features then gathered over NVLink" if gather_input else "; layer 1 gathers K and Q in line"))},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": None if dom is None else {
                "bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                "frac_of_nominal_8TBs": stages[dom]["gbs"] / 8000.0,
                "edge_stage_total": {"ms": tot_ms, "bytes": tot_by, "gbs": tot_by / tot_ms / 1e6,
                                     "frac": tot_by / tot_ms / 1e6 / peak,
                                     "share_of_step": sum(v[0] for v in per.values()) / (ms_step * args.steps)},
                "stages": stages},
            "cpu_baseline": cpu, "parity_check": parity,
            "graph_build_s": t_build, "peak_mem_gib": peak_mem, "phases_ms_per_step": phases,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # libraries (NCCL's version banner, ...) write to fd 1: keep the real stdout for the ONE JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="P", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's nodes/edges (debug)")
    ap.add_argument("--cpu-edges", type=int, default=7_812_500, help="edges of the CPU sample of workload P (1/256)")
    ap.add_argument("--dtype", default=None, choices=sorted(DTYPES), help="table dtype override (A: f32 and bf16 lines)")
    ap.add_argument("--agg", default=None, choices=["sum", "mean", "sym", "max"], help="aggregator override (C: max as published)")
    ap.add_argument("--no-parity-check", action="store_true", help="N>1: skip the pre-timing parity gate")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=4,
                    help="N>1: destination chunks of the cross-layer K prefetch (1 = gather each K table whole)")
    ap.add_argument("--layout-chunks", type=int, default=0,
                    help="N>1: chunk-major layout of the gathered tables (0 = as many as --chunks; 1 = rank-major)")
    ap.add_argument("--bwd-chunks", type=int, default=1,
                    help="N>1: source chunks of the dK walk under which the layer below's dA table travels (1 = off)")
    ap.add_argument("--gather", default="projections", choices=["inputs", "projections"],
                    help="N>1: what travels between layers — the K and Q projections (default) or the layer input "
                         "(every rank then projects the whole K/Q tables locally)")
    ap.add_argument("--no-input-gather", action="store_true",
                    help="N>1: do not gather the node features ahead of the layers (layer 1 gathers K and Q in line)")
    ap.add_argument("--e2e-input-gather", action="store_true",
                    help="N>1: the e2e leg also gathers every step's node features over NVLink behind their H2D copy")
    ap.add_argument("--transport", default="auto", choices=["auto", "push", "pushsm", "pushtma", "peer", "collective"],
                    help="N>1: how row tables travel between ranks (partition.RowPartition.transport)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3
    w = dict(WORKLOADS[args.workload])
    if args.dtype:
        w["dtype"] = args.dtype
    if args.agg:
        w["agg"] = args.agg
    if args.impl == "reference":
        run_reference(args, w)
    elif args.workload == "P":
        run_gpu(args, w)
    else:
        run_small(args, w)


if __name__ == "__main__":
    main()
