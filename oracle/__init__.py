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
