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
