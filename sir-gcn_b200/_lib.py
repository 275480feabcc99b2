"""ctypes binding of libsirgcn.so (include/sirgcn.h).  There is no CPU fallback: if the
library is missing or a call fails, we raise."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SIRGCN_LIB") or os.path.join(_HERE, "libsirgcn.so")   # override: kernel-variant A/B runs

ABI_VERSION = 9     # include/sirgcn.h SIRGCN_ABI_VERSION
WALK_PLAIN_GRID = 1  # sirgcn_edge_args.flags
MAX_ETYPES = 8      # SIRGCN_MAX_ETYPES
F32, BF16, F16 = 0, 1, 2
ACT_IDENTITY, ACT_RELU, ACT_LEAKY_RELU, ACT_GELU = 0, 1, 2, 3

DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

# every symbol include/sirgcn.h declares (tests check the .so exports all of them)
EXPORTS = (
    "sirgcn_last_error", "sirgcn_abi_version", "sirgcn_launch_count",
    "sirgcn_csr_build_workspace_bytes", "sirgcn_csr_build", "sirgcn_schedule_build",
    "sirgcn_edge_subgraph_workspace_bytes", "sirgcn_edge_subgraph",
    "sirgcn_num_tiles", "sirgcn_tiles_build", "sirgcn_rows_build_workspace_bytes", "sirgcn_rows_build",
    "sirgcn_edge_partial_bytes", "sirgcn_edge_fwd", "sirgcn_edge_bwd_q", "sirgcn_edge_bwd_k", "sirgcn_etable_grad",
    "sirgcn_gemm_tn", "sirgcn_gemm_wgrad_workspace_bytes", "sirgcn_gemm_wgrad_plan", "sirgcn_gemm_wgrad", "sirgcn_colsum_workspace_bytes", "sirgcn_colsum", "sirgcn_copy_rows", "sirgcn_mask_scale", "sirgcn_gather_add", "sirgcn_segment_sum", "sirgcn_segment_minmax", "sirgcn_segment_minmax_bwd",
    "sirgcn_peer_alloc", "sirgcn_peer_free", "sirgcn_peer_open", "sirgcn_peer_close", "sirgcn_peer_copy", "sirgcn_peer_push", "sirgcn_peer_push_tma", "sirgcn_peer_barrier",
)


class Schedule(C.Structure):
    _fields_ = [("long_rows", C.c_void_p), ("long_first", C.c_void_p), ("long_nchunks", C.c_void_p),
                ("chunk_lrow", C.c_void_p), ("chunk_beg", C.c_void_p), ("big_lrows", C.c_void_p)]


class EdgeArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32), ("d", C.c_int32), ("dtype", C.c_int32), ("act", C.c_int32),
        ("act_param", C.c_float), ("long_threshold", C.c_int32),
        ("indptr", C.c_void_p), ("idx", C.c_void_p), ("eid", C.c_void_p),
        ("q", C.c_void_p), ("ldq", C.c_int64),
        ("k", C.c_void_p), ("ldk", C.c_int64),
        ("da", C.c_void_p), ("lda", C.c_int64),
        ("e", C.c_void_p), ("lde", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("de", C.c_void_p), ("ldde", C.c_int64),
        ("da_scaled", C.c_void_p), ("ldds", C.c_int64),
        ("dst_scale", C.c_void_p), ("src_scale", C.c_void_p),
        ("sched", Schedule), ("n_long", C.c_int32), ("n_chunks", C.c_int32),
        ("partial", C.c_void_p),
        ("tile_row", C.c_void_p), ("n_tiles", C.c_int32), ("accumulate", C.c_int32),
        ("n_etypes", C.c_int32), ("de_partial", C.c_void_p), ("flags", C.c_int32),
    ]


_lib = None


def lib():
    """Load libsirgcn.so (built by `__graft_entry__.build()` / `make -C sir-gcn_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is required (there is no CPU fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
        l = C.CDLL(LIB_PATH)
        l.sirgcn_last_error.restype = C.c_char_p
        l.sirgcn_launch_count.restype = C.c_uint64
        l.sirgcn_csr_build_workspace_bytes.restype = C.c_size_t
        l.sirgcn_csr_build_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
        l.sirgcn_edge_subgraph_workspace_bytes.restype = C.c_size_t
        l.sirgcn_edge_subgraph_workspace_bytes.argtypes = [C.c_int64]
        l.sirgcn_rows_build_workspace_bytes.restype = C.c_size_t
        l.sirgcn_rows_build_workspace_bytes.argtypes = [C.c_int64, C.c_int32]
        l.sirgcn_colsum_workspace_bytes.restype = C.c_size_t
        l.sirgcn_colsum_workspace_bytes.argtypes = [C.c_int32]
        l.sirgcn_gemm_wgrad_workspace_bytes.restype = C.c_size_t
        l.sirgcn_gemm_wgrad_workspace_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        l.sirgcn_num_tiles.restype = C.c_int64
        l.sirgcn_num_tiles.argtypes = [C.c_int32, C.c_int64]
        l.sirgcn_edge_partial_bytes.restype = C.c_size_t
        l.sirgcn_edge_partial_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        l.sirgcn_abi_version.restype = C.c_int
        if l.sirgcn_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {l.sirgcn_abi_version()}, this package needs {ABI_VERSION}: "
                               "rebuild it (`python -c 'import __graft_entry__ as g; g.build()'`)")
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {lib().sirgcn_last_error().decode()}")


def ptr(t):
    """raw device pointer of a tensor (None -> NULL)"""
    return C.c_void_p(None if t is None else t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch_count() -> int:
    return int(lib().sirgcn_launch_count())
