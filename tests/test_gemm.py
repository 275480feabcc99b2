"""tcgen05/TMEM/TMA projection kernel (sirgcn_gemm_tn) against a plain PyTorch fp32 reference of the same op.
Tolerance: the output is rounded to bf16/fp16 once (rel 2^-8 / 2^-11 of the row maximum) on top of fp32
accumulation of exactly representable 16-bit products, so 1e-2 (bf16) / 2e-3 (fp16) relative to the tensor max."""
import pytest
import torch

import sirgcn_b200  # noqa: F401
from sirgcn_b200 import gemm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return (a.float() - b).abs().max().item() / max(b.abs().max().item(), 1e-20)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize("m,n,k", [(128, 256, 128), (1, 16, 8), (127, 128, 64), (300, 256, 128), (1000, 96, 200),
                                   (4096, 264, 512), (777, 512, 72), (20000, 256, 128), (129, 8, 40)])
def test_gemm_tn_matches_fp32_reference(dtype, tol, m, n, k):
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV).to(dtype)
    b = (torch.randn(n, k, device=DEV) / k ** 0.5).to(dtype)
    bias = torch.randn(n, device=DEV)
    ref = a.float() @ b.float().t()
    out = gemm.gemm_tn(a, b)
    assert out.dtype == dtype and out.shape == (m, n)
    assert rel(out, ref) < tol
    out_b = gemm.gemm_tn(a, b, bias)
    assert rel(out_b, ref + bias) < tol


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (1, 16, 4), (127, 128, 64), (300, 256, 128), (1000, 96, 200),
                                   (4096, 264, 512), (777, 512, 72), (20000, 512, 128), (129, 8, 40), (2944, 152, 64),
                                   (515, 76, 76)])
def test_gemm_tn_fp32_by_3xtf32(m, n, k):
    """fp32 tables on the tensor cores (hi/lo TF32 split, 3 MMAs): against the fp64 product, 1e-5 of the tensor max —
    the fp32 parity target of north_star; a single TF32 product would sit at 1e-3"""
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV)
    b = torch.randn(n, k, device=DEV) / k ** 0.5
    bias = torch.randn(n, device=DEV)
    ref = (a.double() @ b.double().t())
    out = gemm.gemm_tn(a, b)
    assert out.dtype == torch.float32 and out.shape == (m, n)
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 5e-6, err
    out_b = gemm.gemm_tn(a, b, bias)
    assert (out_b.double() - (ref + bias.double())).abs().max().item() / ref.abs().max().item() < 5e-6
    assert torch.equal(out, gemm.gemm_tn(a, b))
    # strided views: one half of a [m, 2n] buffer as output, a padded table as input
    if k % 8 == 0:
        abuf = torch.randn(m, k + 8, device=DEV)
        cbuf = torch.full((m, 2 * n + 8), 7.0, device=DEV)
        gemm.gemm_tn(abuf[:, 4:4 + k], b, out=cbuf[:, 4:4 + n])
        ref2 = abuf[:, 4:4 + k].double() @ b.double().t()
        assert (cbuf[:, 4:4 + n].double() - ref2).abs().max().item() / ref2.abs().max().item() < 5e-6
        assert bool((cbuf[:, :4] == 7).all()) and bool((cbuf[:, 4 + n:] == 7).all())


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-3), (torch.float16, 1e-3)])
@pytest.mark.parametrize("m,n_out,k_in", [(64, 128, 64), (1, 8, 8), (63, 16, 24), (1000, 256, 128), (5000, 152, 72),
                                          (20011, 512, 128), (4097, 512, 256), (300, 1024, 64), (2944, 128, 64),
                                          (100000, 256, 128), (777, 264, 520)])
def test_gemm_wgrad_matches_fp64_reference(dtype, tol, m, n_out, k_in):
    """dW = dY^T·X and db = colsum(dY) by the tcgen05 MN-major kernel: the products of 16-bit values are exact in fp32,
    so the only error is fp32 accumulation over m rows (split over the SMs): relative to the tensor max"""
    torch.manual_seed(m + n_out + k_in)
    dy = torch.randn(m, n_out, device=DEV).to(dtype)
    x = torch.randn(m, k_in, device=DEV).to(dtype)
    assert gemm.wgrad_tc_eligible(dy, x)
    dw, db = gemm.linear_wgrad_bias(dy, x, torch.float32, True)
    ref = dy.double().t() @ x.double()
    refb = dy.double().sum(0)
    assert dw.shape == (n_out, k_in) and db.shape == (n_out,)
    assert (dw.double() - ref).abs().max().item() / ref.abs().max().item() < tol * 1e-2, "dW"
    assert (db.double() - refb).abs().max().item() / max(refb.abs().max().item(), 1.0) < tol * 1e-2, "db"
    dw2, db2 = gemm.linear_wgrad_bias(dy, x, torch.float32, True)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)                # split order is fixed
    dw3, none = gemm.linear_wgrad_bias(dy, x, torch.float32, False)
    assert none is None and (dw3.double() - ref).abs().max().item() / ref.abs().max().item() < tol * 1e-2
    # strided views (the halves of a [m, 2·ld] buffer, padded tables)
    buf = torch.randn(m, n_out + k_in + 16, device=DEV).to(dtype)
    dyv, xv = buf[:, 8:8 + n_out], buf[:, 8 + n_out:8 + n_out + k_in]
    dwv, dbv = gemm.linear_wgrad_bias(dyv, xv, torch.float32, True)
    refv = dyv.double().t() @ xv.double()
    assert (dwv.double() - refv).abs().max().item() / refv.abs().max().item() < tol * 1e-2
    assert (dbv.double() - dyv.double().sum(0)).abs().max().item() / max(dyv.double().sum(0).abs().max().item(), 1.0) < tol * 1e-2


def test_gemm_tn_strided_operands_and_output_view():
    """Q|K halves and padded tables are strided views: ld > row length"""
    torch.manual_seed(0)
    m, n, k = 515, 80, 72
    abuf = torch.randn(m, 96, device=DEV).to(torch.bfloat16)
    a = abuf[:, :k]
    b = (torch.randn(n, k, device=DEV) / 8).to(torch.bfloat16)
    cbuf = torch.full((m, 128), 7.0, device=DEV, dtype=torch.bfloat16)
    gemm.gemm_tn(a, b, out=cbuf[:, 8:8 + n])
    ref = a.float() @ b.float().t()
    assert rel(cbuf[:, 8:8 + n], ref) < 1e-2
    assert bool((cbuf[:, :8] == 7).all()) and bool((cbuf[:, 8 + n:] == 7).all())      # nothing outside the view


def test_gemm_repeatable_and_layer_uses_it():
    from torch import nn
    from sirgcn_b200 import Graph, SIRConv, _lib
    torch.manual_seed(0)
    a = torch.randn(5000, 128, device=DEV).to(torch.bfloat16)
    b = torch.randn(256, 128, device=DEV).to(torch.bfloat16)
    assert torch.equal(gemm.gemm_tn(a, b), gemm.gemm_tn(a, b))
    src, dst = torch.randint(0, 500, (4000,), device=DEV), torch.randint(0, 500, (4000,), device=DEV)
    g = Graph(src, dst, 500)
    layer = SIRConv(64, 128, 64, nn.ReLU(), agg_type="mean").to(DEV)
    x = torch.randn(500, 64, device=DEV).to(torch.bfloat16).requires_grad_(True)
    before = _lib.launch_count()
    out = layer(g, x)
    out.backward(torch.randn_like(out))
    # 2 forward GEMMs + 2 dgrads + 3 edge passes at least
    assert _lib.launch_count() - before >= 7
    ref = SIRConv(64, 128, 64, nn.ReLU(), agg_type="mean").to(DEV)
    ref.load_state_dict(layer.state_dict())
    out32 = ref(g, x.detach().float())
    assert rel(out, out32.detach()) < 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n", [(1, 8), (1000, 128), (70001, 256), (333, 96), (5, 1024)])
def test_column_sum(dtype, m, n):
    """bias-gradient reduction (sirgcn_colsum): fp32 accumulation of the stored values, bitwise repeatable"""
    if dtype != torch.float32 and n == 1024 and False:
        pytest.skip()
    torch.manual_seed(m + n)
    x = torch.randn(m, n, device=DEV).to(dtype)
    got = gemm.column_sum(x)
    ref = x.double().sum(0)
    assert got.dtype == torch.float32 and got.shape == (n,)
    assert (got.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()) * max(1.0, m ** 0.5 / 30)
    assert torch.equal(got, gemm.column_sum(x))
    view = torch.randn(m, n + 16, device=DEV).to(dtype)[:, 8:8 + n] if dtype != torch.float32 else None
    if view is not None:      # strided 16-byte aligned view
        assert (gemm.column_sum(view).double() - view.double().sum(0)).abs().max().item() <= 1e-4 * max(1.0, m ** 0.5)
