"""The C-ABI library loads on a machine without a GPU and exports exactly the symbols that
include/sirgcn.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sirgcn.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sirgcn_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    import sirgcn_b200
    from sirgcn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.LIB_PATH


def test_header_matches_binding_table():
    from sirgcn_b200 import _lib
    assert declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in sirgcn.h but not exported"
    lib.sirgcn_abi_version.restype = ctypes.c_int
    assert lib.sirgcn_abi_version() == 9
    lib.sirgcn_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.sirgcn_last_error(), bytes)


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_torch_types_in_abi():
    text = open(HEADER).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text


def test_argument_validation_without_gpu(lib_path):
    """bad arguments are rejected before any CUDA call, with a message"""
    from sirgcn_b200 import _lib
    L = _lib.lib()
    a = _lib.EdgeArgs()
    a.n_rows, a.d, a.dtype, a.act, a.long_threshold = 4, 0, 0, 0, 512
    assert L.sirgcn_edge_fwd(ctypes.byref(a), None) == -1
    assert b"d=0" in L.sirgcn_last_error()
    a.d, a.dtype = 8, 7
    assert L.sirgcn_edge_fwd(ctypes.byref(a), None) == -1
    a.dtype, a.long_threshold = 0, 1
    assert L.sirgcn_edge_bwd_k(ctypes.byref(a), None) == -1
    with pytest.raises(RuntimeError, match="long_threshold"):
        _lib.check(-1, "sirgcn_edge_bwd_k")


def test_wgrad_launch_plan_invariants(lib_path):
    """sirgcn_gemm_wgrad's host-side plan (no GPU needed): every 64-node block is covered by exactly one split, a CTA
    never accumulates more than 512 blocks (the tensor core's accumulation truncates: DESIGN §2.4), the accumulators of
    all M tiles fit the 512 TMEM columns, the ring fits shared memory, and the workspace query matches the plan"""
    lib = ctypes.CDLL(lib_path)
    lib.sirgcn_gemm_wgrad_workspace_bytes.restype = ctypes.c_size_t
    lib.sirgcn_gemm_wgrad_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32]
    plan = (ctypes.c_int32 * 10)()
    for m in (1, 63, 64, 65, 2944, 100_000, 4_850_000, 6_250_000, 25_000_000, 50_000_000, 2_000_000_000):
        for n_out, k_in in ((8, 8), (128, 64), (152, 72), (256, 128), (512, 128), (512, 256), (1024, 64), (264, 520)):
            for bias in (0, 1):
                assert lib.sirgcn_gemm_wgrad_plan(ctypes.c_int64(m), n_out, k_in, bias, plan) == 0
                splits, per, n_ranges, m_groups, mt, nc, ncp, stages, tmem, smem = list(plan)
                blocks = (m + 63) // 64
                assert splits >= 1 and 1 <= per <= 512, (m, n_out, k_in, per)
                assert (splits - 1) * per < blocks <= splits * per, (m, blocks, splits, per)     # no empty split
                assert 1 <= mt <= 4 and m_groups * mt * 128 >= n_out and n_ranges * nc >= k_in
                assert nc % 16 == 0 and 16 <= nc <= 256 and ncp == nc + 16 * bias
                assert mt * ncp + 16 <= tmem <= 512 and tmem & (tmem - 1) == 0
                assert 2 <= stages <= 8 and smem <= 227 * 1024
                need = lib.sirgcn_gemm_wgrad_workspace_bytes(ctypes.c_int64(m), n_out, k_in)
                ldp = (k_in + 3) // 4 * 4 + 4
                assert need >= splits * n_out * ldp * 4
