#!/usr/bin/env python
"""Row-table all-gather bandwidth of the transports of sir-gcn_b200/peer.py, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/peer_bw.py

Each rank owns a [rows, 128] bf16 slice (default 6.25 M rows = the K slice of config P at 8 GPUs) and receives the
other ranks' slices.  Prints one JSON line: GB/s received per rank (max time over ranks, CUDA events) per variant."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=6_250_000)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import sirgcn_b200  # noqa: F401
    from sirgcn_b200 import peer
    pool = peer.PeerPool()
    rows, ld, dt = args.rows, 128, torch.bfloat16
    sl = pool.acquire(rows, ld, dt)
    sl.local.copy_(torch.full((rows, ld), float(rank + 1), dtype=dt, device=dev))
    src = sl.local.clone()
    full = torch.empty((world * rows, ld), dtype=dt, device=dev)
    pf = pool.acquire_full(rows, ld, dt)
    nbytes = (world - 1) * rows * ld * 2

    def check(t):
        got = t.view(world, rows, ld)[:, ::max(1, rows // 64), 0].float().mean(1).cpu()
        assert torch.equal(got, torch.arange(1, world + 1, dtype=torch.float32)), got

    def run(fn, target):
        target.zero_()
        fn()
        torch.cuda.synchronize()
        check(target)
        times = []
        for _ in range(args.iters):
            dist.barrier()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            fn()
            t1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            times.append(float(ms))
        best = min(times)
        return {"ms": round(best, 3), "gbs_in_per_rank": round(nbytes / best / 1e6, 1)}

    def pull():
        pool.ensure_writable(sl)
        pool.gather(sl, full).wait()

    def push(mode):
        def f():
            pool.push(src, pf, mode=mode).wait()
            pf.released_at = pool.barriers
        return f

    def nccl():
        dist.all_gather_into_tensor(full, src)

    res = {"world": world, "slice_mb": rows * ld * 2 / 1e6}
    res["nccl_all_gather"] = run(nccl, full)
    res["ce_pull"] = run(pull, full)
    res["ce_push"] = run(push("ce"), pf.local)
    for ctas in (8, 16, 32, 64, 128):
        pool.push_ctas = ctas
        res[f"sm_push_{ctas}ctas"] = run(push("sm"), pf.local)
    for ctas in (4, 8, 16, 32):
        pool.push_ctas = ctas
        res[f"tma_push_{ctas}ctas"] = run(push("tma"), pf.local)
    pool.check()
    if rank == 0:
        print(json.dumps(res), flush=True)
    pool.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
