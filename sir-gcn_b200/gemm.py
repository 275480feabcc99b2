"""Dense projections of the layer (linear_query ‖ linear_key, linear_relation;
/root/reference/models/conv.py:60-61,:65).

bf16 operands on CUDA go to the hand-written tcgen05/TMEM/TMA kernel (`sirgcn_gemm_tn`); fp32
operands keep ATen's SGEMM (TF32 off, as in the reference) because tcgen05 has no IEEE-fp32 MMA
and the parity target for fp32 is 1e-5 relative.
"""
from __future__ import annotations

import torch.nn.functional as F


def linear(x, weight, bias=None):
    return F.linear(x, weight, bias)
