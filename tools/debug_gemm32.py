import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sirgcn_b200
from sirgcn_b200 import gemm
dev = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False
def relmax(a, b): return float((a.double() - b).abs().max() / b.abs().max())
def relrms(a, b): return float((a.double() - b).norm() / b.norm())
for (m, n, k, scale) in [(2944, 128, 64, 1.0), (15000, 256, 128, 1.0), (15000, 128, 128, 30.0), (169343, 512, 128, 1.0), (15000, 128, 256, 1.0)]:
    torch.manual_seed(0)
    a = torch.randn(m, k, device=dev) * scale
    b = torch.randn(n, k, device=dev) / k ** 0.5
    ref = a.double() @ b.double().t()
    tc = gemm.gemm_tn(a, b)
    sg = a @ b.t()
    print(f"m={m} n={n} k={k} scale={scale}: tf32x4 max {relmax(tc, ref):.2e} rms {relrms(tc, ref):.2e} | sgemm max {relmax(sg, ref):.2e} rms {relrms(sg, ref):.2e}", flush=True)
    # structured data: a column with a large constant offset (cancellation)
    a2 = a + 100.0
    ref2 = a2.double() @ b.double().t()
    print(f"    offset+100: tf32x4 max {relmax(gemm.gemm_tn(a2, b), ref2):.2e} | sgemm {relmax(a2 @ b.t(), ref2):.2e}", flush=True)
