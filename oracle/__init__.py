"""Parity oracle (test infrastructure).  See sirconv_ref.py / csr_ref.c headers."""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcsr_ref.so")


def build_c_oracle():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def csr_ref_c(src, dst, num_nodes):
    """Run oracle/csr_ref.c on int32 COO tensors; same outputs as sirconv_ref.csr_csc_ref."""
    if not os.path.exists(_SO):
        build_c_oracle()
    lib = ctypes.CDLL(_SO)
    src = torch.as_tensor(src).to(torch.int32).contiguous()
    dst = torch.as_tensor(dst).to(torch.int32).contiguous()
    E, N = src.numel(), int(num_nodes)
    i32 = lambda n: torch.empty(n, dtype=torch.int32)
    outs = [i32(N + 1), i32(E), i32(E), i32(N + 1), i32(E), i32(E),
            torch.empty(N, dtype=torch.float32), torch.empty(N, dtype=torch.float32)]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.csr_ref_build(p(src), p(dst), ctypes.c_int64(E), ctypes.c_int32(N), *[p(t) for t in outs])
    if rc:
        raise ValueError(f"csr_ref_build failed rc={rc}")
    return tuple(outs)


_SO2 = os.path.join(_HERE, "_build", "libsirconv_ref.so")
_ACT = {"identity": 0, "relu": 1, "leaky": 2, "gelu": 3}
_AGG = {"sum": 0, "mean": 1, "sym": 2}


def edge_stage_c(src, dst, num_nodes, q, k, e, act, slope, agg, da=None):
    """oracle/sirconv_ref.c (plain C loops, fp64): A = edge stage forward, and with `da` also (dQ, dK, dE)"""
    if not os.path.exists(_SO2):
        build_c_oracle()
    lib = ctypes.CDLL(_SO2)
    src = torch.as_tensor(src).to(torch.int64).contiguous()
    dst = torch.as_tensor(dst).to(torch.int64).contiguous()
    q, k = q.detach().double().contiguous(), k.detach().double().contiguous()
    e = None if e is None else e.detach().double().contiguous()
    E, N, d = src.numel(), int(num_nodes), q.shape[1]
    p = lambda t: ctypes.c_void_p(None if t is None else t.data_ptr())
    head = (ctypes.c_int64(E), ctypes.c_int32(N), ctypes.c_int32(d), p(src), p(dst), p(q), p(k), p(e))
    tail = (ctypes.c_int(_ACT[act]), ctypes.c_double(slope), ctypes.c_int(_AGG[agg]))
    a = torch.empty(N, d, dtype=torch.float64)
    rc = lib.sirconv_ref_edge_forward(*head, *tail, p(a))
    if rc:
        raise ValueError(f"sirconv_ref_edge_forward failed rc={rc}")
    if da is None:
        return a
    da = da.detach().double().contiguous()
    dq, dk = torch.empty(N, d, dtype=torch.float64), torch.empty(N, d, dtype=torch.float64)
    de = None if e is None else torch.empty(E, d, dtype=torch.float64)
    rc = lib.sirconv_ref_edge_backward(*head, p(da), *tail, p(dq), p(dk), p(de))
    if rc:
        raise ValueError(f"sirconv_ref_edge_backward failed rc={rc}")
    return a, dq, dk, de
