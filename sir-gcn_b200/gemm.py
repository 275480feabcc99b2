"""Dense projections of the layer (linear_query ‖ linear_key, linear_relation;
/root/reference/models/conv.py:60-61,:65) and their backward.

All three GEMM shapes are "TN" products of row-major operands:
    forward   C[M, N]  = X[M, K] · W[N, K]^T (+ b)
    dgrad     dX[M, K] = dY[M, N] · W[N, K]          = dY · (W^T)^T
    wgrad     dW[N, K] = dY[M, N]^T · X[M, K]
bf16 operands on CUDA go to the hand-written tcgen05/TMEM/TMA kernel when it is built in;
fp32 operands keep ATen's SGEMM (TF32 off, as in the reference) because tcgen05 has no
IEEE-fp32 MMA and the fp32 parity target is 1e-5 relative.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def linear(x, weight, bias=None):
    """autograd-visible projection used by the composed (dropout / split) paths"""
    return F.linear(x, weight, bias)


def linear_forward(x, weight, bias):
    """no-autograd forward used inside SIRLayerFunction; output dtype follows autocast / x"""
    if torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
        x, weight = x.to(dt), weight.to(dt)
        bias = None if bias is None else bias.to(dt)
    elif weight.dtype != x.dtype:
        weight = weight.to(x.dtype)
        bias = None if bias is None else bias.to(x.dtype)
    return F.linear(x, weight, bias)


def linear_dgrad(dy, weight, pad_to=None):
    """dX = dY · W; with pad_to, the result is a [M, pad_to] buffer whose extra columns are zero"""
    k = weight.shape[1]
    if pad_to is None or pad_to == k:
        return dy @ weight
    out = torch.zeros((dy.shape[0], pad_to), dtype=dy.dtype, device=dy.device)
    out[:, :k].copy_(dy @ weight)
    return out


def linear_wgrad(dy, x, out_dtype):
    """dW = dY^T · X, accumulated in fp32 by the GEMM, returned in the parameter dtype"""
    return (dy.t() @ x.to(dy.dtype)).to(out_dtype)
