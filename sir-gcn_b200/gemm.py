"""Dense projections of the layer (linear_query ‖ linear_key, linear_relation;
/root/reference/models/conv.py:60-61,:65) and their backward.

All three GEMM shapes are "TN" products of row-major operands:
    forward   C[M, N]  = X[M, K] · W[N, K]^T (+ b)
    dgrad     dX[M, K] = dY[M, N] · W[N, K]          = dY · (W^T)^T
    wgrad     dW[N, K] = dY[M, N]^T · X[M, K]
16-bit operands on CUDA go to the hand-written tcgen05/TMEM/TMA kernel; fp32 operands go to its 3xTF32 variant
(hi/lo split in shared memory, three kind::tf32 MMAs per K step, fp32 accumulation in TMEM: ~1e-6 relative, inside
the 1e-5 fp32 parity target that a single TF32 product misses — tcgen05 has no IEEE-fp32 MMA).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn.functional as F

from . import _lib

# SIRGCN_GEMM=cublas routes the projections through the library instead (A/B measurements); SIRGCN_GEMM32=cublas
# only the fp32 ones
_USE_TC = os.environ.get("SIRGCN_GEMM", "tcgen05") != "cublas"
_USE_TC32 = _USE_TC and os.environ.get("SIRGCN_GEMM32", "tcgen05") != "cublas"


def _rows16(t):
    """2-D, unit column stride, 16-byte aligned rows"""
    return (t.dim() == 2 and t.stride(1) == 1 and t.data_ptr() % 16 == 0 and
            (t.shape[0] <= 1 or (t.stride(0) * t.element_size()) % 16 == 0))


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def tc_eligible(x, n, k):
    """the hand-written tcgen05 kernels handle CUDA bf16 / fp16 / fp32 operands whose rows are whole 16-byte vectors
    (n, k multiples of 8 for 16-bit, of 4 for fp32)"""
    if not x.is_cuda or x.shape[0] >= 2 ** 31:
        return False
    if x.dtype in (torch.bfloat16, torch.float16):
        return _USE_TC and n % 8 == 0 and k % 8 == 0
    if x.dtype == torch.float32:
        return _USE_TC32 and n % 4 == 0 and k % 4 == 0
    return False


def gemm_tn(a, b, bias=None, out=None):
    """C[m, n] = a[m, k] · b[n, k]^T (+ bias) on the tcgen05 tensor cores (sirgcn_gemm_tn, include/sirgcn.h)"""
    m, k = a.shape
    n = b.shape[0]
    if not _rows16(a):
        a = a.contiguous()
    if not _rows16(b):
        b = b.contiguous()
    if out is None:
        out = torch.empty((m, n), dtype=a.dtype, device=a.device)
    if bias is not None:
        bias = bias.detach().to(torch.float32).contiguous()
    with torch.cuda.device(a.device):
        rc = _lib.lib().sirgcn_gemm_tn(_lib.ptr(a), C.c_int64(_ld(a)), _lib.ptr(b), C.c_int64(_ld(b)), _lib.ptr(out),
                                       C.c_int64(_ld(out)), _lib.ptr(bias), C.c_int64(m), C.c_int32(n), C.c_int32(k),
                                       C.c_int32(_lib.DTYPE_CODE[a.dtype]), _lib.stream_ptr(a.device))
    _lib.check(rc, "sirgcn_gemm_tn")
    return out


def column_sum(x, out_dtype=torch.float32):
    """Σ_rows x in fp32 (bias gradients): sirgcn_colsum on CUDA tables with 16-byte rows, torch elsewhere"""
    es = x.element_size()
    n = x.shape[1]
    if (x.is_cuda and x.dtype in _lib.DTYPE_CODE and x.dim() == 2 and x.stride(1) == 1 and x.data_ptr() % 16 == 0
            and (n * es) % 16 == 0 and (n * es) // 16 <= 256 and (x.shape[0] <= 1 or (x.stride(0) * es) % 16 == 0)):
        L = _lib.lib()
        out = torch.empty(n, dtype=torch.float32, device=x.device)
        nbytes = L.sirgcn_colsum_workspace_bytes(C.c_int32(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rc = L.sirgcn_colsum(_lib.ptr(x), C.c_int64(_ld(x)), C.c_int64(x.shape[0]), C.c_int32(n),
                                 C.c_int32(_lib.DTYPE_CODE[x.dtype]), _lib.ptr(out), _lib.ptr(ws), C.c_size_t(nbytes),
                                 _lib.stream_ptr(x.device))
        _lib.check(rc, "sirgcn_colsum")
        return out.to(out_dtype)
    return x.sum(0, dtype=torch.float64 if x.dtype == torch.float64 else torch.float32).to(out_dtype)


def linear(x, weight, bias=None):
    """autograd-visible projection used by the composed (dropout / split) paths"""
    return F.linear(x, weight, bias)


def linear_forward(x, weight, bias, out=None):
    """no-autograd forward used inside SIRLayerFunction; output dtype follows autocast / x"""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        x, weight = x.to(dt), weight.to(dt)
        bias = None if bias is None else bias.to(dt)
    elif weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
    if x.dim() == 2 and tc_eligible(x, weight.shape[0], weight.shape[1]):
        return gemm_tn(x, weight.detach(), bias, out=out)
    res = F.linear(x, weight, None if bias is None else bias.to(x.dtype))
    return res if out is None else out.copy_(res)


def linear_dgrad(dy, weight, pad_to=None, out=None):
    """dX = dY · W; with pad_to, the result is a [M, pad_to] buffer whose extra columns are zero"""
    k = weight.shape[1]
    if out is not None:
        if dy.dim() == 2 and tc_eligible(dy, k, weight.shape[0]):
            return gemm_tn(dy, weight.detach().t().contiguous(), out=out)
        return out.copy_(dy @ weight)
    if dy.dim() == 2 and tc_eligible(dy, k, weight.shape[0]):
        wt = weight.detach().t().contiguous()               # [in, out]: K-major B operand of the TN kernel (tiny)
        if pad_to is None or pad_to == k:
            return gemm_tn(dy, wt)
        out = torch.zeros((dy.shape[0], pad_to), dtype=dy.dtype, device=dy.device)
        gemm_tn(dy, wt, out=out[:, :k])
        return out
    if pad_to is None or pad_to == k:
        return dy @ weight
    out = torch.zeros((dy.shape[0], pad_to), dtype=dy.dtype, device=dy.device)
    out[:, :k].copy_(dy @ weight)
    return out


def wgrad_tc_eligible(dy, x):
    return (_USE_TC and os.environ.get("SIRGCN_WGRAD", "tcgen05") != "cublas" and dy.is_cuda and x.is_cuda
            and dy.dtype in (torch.bfloat16, torch.float16) and x.dtype == dy.dtype and dy.dim() == 2 and x.dim() == 2
            and dy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0 and dy.shape[0] < 2 ** 31 and _rows16(dy) and _rows16(x))


def linear_wgrad_bias(dy, x, out_dtype, want_bias):
    """(dW = dY^T · X, db = column sums of dY or None): for 16-bit tables ONE pass of the hand-written tcgen05 kernel
    over both tables (sirgcn_gemm_wgrad: MN-major operands, node dimension split over the SMs, bias as an extra MMA
    against a tile of ones); otherwise the library GEMM + sirgcn_colsum.  fp32 accumulation either way."""
    if wgrad_tc_eligible(dy, x):
        m, n_out, k_in = dy.shape[0], dy.shape[1], x.shape[1]
        L = _lib.lib()
        dw = torch.empty((n_out, k_in), dtype=torch.float32, device=dy.device)
        db = torch.empty(n_out, dtype=torch.float32, device=dy.device) if want_bias else None
        nbytes = L.sirgcn_gemm_wgrad_workspace_bytes(C.c_int64(m), C.c_int32(n_out), C.c_int32(k_in))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dy.device)
        with torch.cuda.device(dy.device):
            rc = L.sirgcn_gemm_wgrad(_lib.ptr(dy), C.c_int64(_ld(dy)), _lib.ptr(x), C.c_int64(_ld(x)), C.c_int64(m),
                                     C.c_int32(n_out), C.c_int32(k_in), C.c_int32(_lib.DTYPE_CODE[dy.dtype]),
                                     _lib.ptr(dw), C.c_int64(k_in), _lib.ptr(db), _lib.ptr(ws), C.c_size_t(nbytes),
                                     _lib.stream_ptr(dy.device))
        _lib.check(rc, "sirgcn_gemm_wgrad")
        return dw.to(out_dtype), (None if db is None else db.to(out_dtype))
    dw = (dy.t() @ x.to(dy.dtype)).to(out_dtype)
    return dw, (column_sum(dy, out_dtype) if want_bias else None)


def linear_wgrad(dy, x, out_dtype):
    """dW = dY^T · X, accumulated in fp32 by the GEMM, returned in the parameter dtype"""
    return linear_wgrad_bias(dy, x, out_dtype, False)[0]
