"""bench.py contract checks that need no GPU: the reference arm's JSON line, and the algorithmic-byte model of
SURVEY.md §8(d) / BASELINE.md §3 that the roofline fractions are computed from."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-edges", "50000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "sirgcn_conv_fwd_bwd_gedges_per_s" and j["unit"] == "Gedges/s"
    assert j["higher_is_better"] is True and j["vs_baseline"] is None and j["gpu_launches"] == 0
    assert j["value"] > 0 and j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["kind"] == "port"
    assert j["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "Gedges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("P:") and j["config"]["edges"] == 2_000_000_000


def test_algorithmic_bytes_match_the_survey():
    import bench
    E, N, r = 2_000_000_000, 50_000_000, 128 * 2               # config P, bf16
    per_layer = sum(bench.edge_bytes(k, E, N, r) for k in ("sirgcn_edge_fwd", "sirgcn_edge_bwd_q", "sirgcn_edge_bwd_k"))
    assert per_layer == E * (4 * r + 12) + N * (7 * r + 12)     # total fwd+bwd, ε = 0
    assert round(per_layer / E) == 1081                         # 1,081 B/edge, 2.16 TB per layer
    assert bench.edge_bytes("sirgcn_edge_fwd", E, N, r) == E * (r + 4) + N * (2 * r + 4) == 545_800_000_000
    # arxiv-shaped, fp32, d = 256: 5,151 B/edge
    E, N, r = 1_166_243, 169_343, 256 * 4
    assert round((E * (4 * r + 12) + N * (7 * r + 12)) / E) == 5151
    # with a projected edge term read by all three walks (ε = r): ZINC-shaped fp32 d = 64 -> 2,634 B/edge
    E, N, r = 6_400, 2_944, 64 * 4
    tot = sum(bench.edge_bytes(k, E, N, r, eps=r) for k in ("sirgcn_edge_fwd", "sirgcn_edge_bwd_q", "sirgcn_edge_bwd_k"))
    assert round(tot / E) == 2634


def _verdict_worker(rank, world, port, ret):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        ok = {"out": 0.0, "dX": 4.5e-3, "dW": 5.3e-3}
        # only rank 0 holds the oracle leg; in case 2 it alone sees a failing number
        c1 = bench.parity_verdict(dict(ok, **({"oracle_out": 3.9e-3} if rank == 0 else {})),
                                  {"oracle_dX_frobenius": 8e-3} if rank == 0 else {}, 2e-2, 5e-2, "cpu")
        c2 = bench.parity_verdict(dict(ok, **({"oracle_out": 6e-2} if rank == 0 else {})), {}, 2e-2, 5e-2, "cpu")
        c3 = bench.parity_verdict(dict(ok, dX=(3e-2 if rank == world - 1 else 1e-3)), {}, 2e-2, 5e-2, "cpu")
        c4 = bench.parity_verdict(ok, {"oracle_dX_frobenius": 0.2} if rank == 0 else {}, 2e-2, 5e-2, "cpu")
        ret[rank] = (c1, c2, c3, c4)
    finally:
        dist.destroy_process_group()


def test_parity_gate_verdict_is_the_same_on_every_rank():
    """the N>1 parity gate of bench.py: whatever a single rank measured, ALL ranks get the same verdict"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 3
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_verdict_worker, args=(world, port, ret), nprocs=world, join=True)
        got = dict(ret)
    assert len(got) == world and all(got[r] == got[0] for r in got)
    c1, c2, c3, c4 = got[0]
    assert c1 == (5.3e-3, True)
    assert c2[1] is False and abs(c2[0] - 6e-2) < 1e-12
    assert c3[1] is False and abs(c3[0] - 3e-2) < 1e-12
    assert c4[1] is False and abs(c4[0] - 5.3e-3) < 1e-12          # failed on the Frobenius leg alone
