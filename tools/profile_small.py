"""host-side phase timings of one small-workload step (Z / C): where do the milliseconds go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
import bench
import sirgcn_b200
from sirgcn_b200 import Graph, SIRConv, SIREConv, DropEdge, _lib

name = sys.argv[1] if len(sys.argv) > 1 else "Z"
w = dict(bench.WORKLOADS[name])
dev = torch.device("cuda:0")
layers = bench.small_layers(w, SIRConv, SIREConv, w.get("edge_types")).to(dev)
b = bench.small_batch(name, w, 0)
b["gout"] = torch.randn(b["n"], w["d"])
b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in b.items()}
params = list(layers.parameters())
def sync(): torch.cuda.synchronize()
def phase(label, fn, reps=20):
    for _ in range(3): fn()
    sync(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    t1 = time.perf_counter(); sync(); t2 = time.perf_counter()
    print(f"{label:32s} host {1e3*(t1-t0)/reps:8.3f} ms/iter   host+drain {1e3*(t2-t0)/reps:8.3f} ms/iter", flush=True)
need = b["etype"] is not None
phase("Graph(src,dst,n)", lambda: Graph(b["src"], b["dst"], b["n"], need_eid=need))
g = Graph(b["src"], b["dst"], b["n"], need_eid=need)
def fwd():
    h = b["x"]
    for layer in layers:
        h = layer(g, h, b["etype"]) if need else layer(g, h)
    return h
with torch.no_grad():
    phase("forward (no_grad)", fwd)
phase("forward (autograd)", fwd)
def fb():
    for p in params: p.grad = None
    fwd().backward(b["gout"])
phase("forward+backward", fb)
def full():
    global g
    g = Graph(b["src"], b["dst"], b["n"], need_eid=need)
    fb()
phase("graph+forward+backward", full)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): full()
    sync()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40))
