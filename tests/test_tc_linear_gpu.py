"""tcgen05 dense kernels against torch fp32 (float reference for a floating-point kernel)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = {3: dict(rtol=2e-4, atol=2e-4), 1: dict(rtol=3e-2, atol=3e-2)}


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


@pytest.mark.parametrize("precision", [3, 1])
@pytest.mark.parametrize("n", [128, 1000])
def test_in_projection_with_position_prologue(precision, n):
    from geomae_b200.dense import tc_linear
    x, W, b = rnd(n, 128, seed=1), rnd(384, 128, scale=0.1, seed=2), rnd(384, seed=3)
    table = rnd(144, 128, seed=4)
    cell = torch.randint(0, 144, (n,), dtype=torch.int32, device="cuda")
    got = tc_linear(x, W, n_out=384, bias=b, pos_table=table, tok_cell=cell, pos_slabs=2, precision=precision)
    xp = x + table[cell.long()]
    ref = torch.cat([F.linear(xp, W[:256], b[:256]), F.linear(x, W[256:], b[256:])], dim=1)
    torch.testing.assert_close(got, ref, **TOL[precision])


@pytest.mark.parametrize("precision", [3, 1])
@pytest.mark.parametrize("K,gelu", [(128, False), (256, True)])
def test_projection_residual_layernorm_epilogue(precision, K, gelu):
    from geomae_b200.dense import tc_linear
    n = 777
    a, W, b, res = rnd(n, K, seed=5), rnd(128, K, scale=0.1, seed=6), rnd(128, seed=7), rnd(n, 128, seed=8)
    gamma, beta = 1 + 0.1 * rnd(128, seed=9), 0.1 * rnd(128, seed=10)
    out, ln_in, stats = tc_linear(a, W, n_out=128, bias=b, add_src=res, a_gelu=gelu, ln=(gamma, beta, 1e-5, True),
                                  precision=precision)
    s = F.linear(F.gelu(a) if gelu else a, W, b) + res
    ref = F.layer_norm(s, (128,), gamma, beta, 1e-5)
    torch.testing.assert_close(ln_in, s, **TOL[precision])
    torch.testing.assert_close(out, ref, **TOL[precision])
    torch.testing.assert_close(stats[:, 0], s.mean(dim=1), **TOL[precision])
    torch.testing.assert_close(stats[:, 1], torch.rsqrt(s.var(dim=1, unbiased=False) + 1e-5), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("precision", [3, 1])
def test_ffn1_plain_256(precision):
    from geomae_b200.dense import tc_linear
    y, W, b = rnd(515, 128, seed=11), rnd(256, 128, scale=0.1, seed=12), rnd(256, seed=13)
    torch.testing.assert_close(tc_linear(y, W, n_out=256, bias=b, precision=precision), F.linear(y, W, b), **TOL[precision])


@pytest.mark.parametrize("precision", [3, 1])
def test_input_gradient_forms(precision):
    from geomae_b200.dense import tc_linear
    n = 900
    # dh = ds2 @ W2, times gelu'(u)          (W2: [128 out, 256 in])
    ds2, W2, u = rnd(n, 128, seed=14), rnd(128, 256, scale=0.1, seed=15), rnd(n, 256, seed=16)
    got = tc_linear(ds2, W2, n_out=256, w_mn_major=True, gelu_u=u, precision=precision)
    uu = u.clone().requires_grad_(True)
    F.gelu(uu).backward(ds2 @ W2)
    torch.testing.assert_close(got, uu.grad, **TOL[precision])
    # dy = ds2 + du @ W1                      (W1: [256 out, 128 in], K = 256)
    du, W1 = rnd(n, 256, seed=17), rnd(256, 128, scale=0.1, seed=18)
    got = tc_linear(du, W1, n_out=128, w_mn_major=True, add_src=ds2, precision=precision)
    torch.testing.assert_close(got, ds2 + du @ W1, **TOL[precision])
    # dx = ds1 + dqkv @ Wqkv                  (Wqkv: [384 out, 128 in], K = 384)
    dqkv, Wqkv = rnd(n, 384, seed=19), rnd(384, 128, scale=0.1, seed=20)
    got = tc_linear(dqkv, Wqkv, n_out=128, w_mn_major=True, add_src=ds2, precision=precision)
    torch.testing.assert_close(got, ds2 + dqkv @ Wqkv, **TOL[precision])


@pytest.mark.parametrize("precision", [3, 1])
def test_weight_gradient_forms(precision):
    from geomae_b200.dense import tc_wgrad
    for n in (100, 1000, 5000):
        # bf16 rounding of ~n unit-variance products: error ~ 2^-9 * sqrt(n) * few sigma
        tol = dict(rtol=1e-3, atol=2e-3) if precision == 3 else dict(rtol=5e-2, atol=0.02 * n ** 0.5)
        # dW2 [128, 256] += ds2^T gelu(u), db2
        ds2, u = rnd(n, 128, seed=21), rnd(n, 256, seed=22)
        dW, db = torch.full((128, 256), 0.5, device="cuda"), torch.full((128,), 0.25, device="cuda")
        tc_wgrad(ds2, u, dW, db, x_gelu=True, precision=precision)
        torch.testing.assert_close(dW, 0.5 + ds2.t() @ F.gelu(u), **tol)
        torch.testing.assert_close(db, 0.25 + ds2.sum(0), **tol)
        # dW1 [256, 128] += du^T y
        du, y = rnd(n, 256, seed=23), rnd(n, 128, seed=24)
        dW, db = torch.zeros(256, 128, device="cuda"), torch.zeros(256, device="cuda")
        tc_wgrad(du, y, dW, db, precision=precision)
        torch.testing.assert_close(dW, du.t() @ y, **tol)
        torch.testing.assert_close(db, du.sum(0), **tol)
        # dWqkv [384, 128] += dqkv^T [x+pos | x+pos | x]
        dqkv, x, table = rnd(n, 384, seed=25), rnd(n, 128, seed=26), rnd(144, 128, seed=27)
        cell = torch.randint(0, 144, (n,), dtype=torch.int32, device="cuda")
        dW, db = torch.zeros(384, 128, device="cuda"), torch.zeros(384, device="cuda")
        tc_wgrad(dqkv, x, dW, db, pos_table=table, tok_cell=cell, pos_slabs=2, precision=precision)
        xp = x + table[cell.long()]
        ref = torch.cat([dqkv[:, :256].t() @ xp, dqkv[:, 256:].t() @ x], dim=0)
        torch.testing.assert_close(dW, ref, **tol)
        torch.testing.assert_close(db, dqkv.sum(0), **tol)


def test_layernorm_backward():
    from geomae_b200.dense import layernorm_bwd
    n = 3001
    s = rnd(n, 128, seed=30).requires_grad_(True)
    gamma = (1 + 0.1 * rnd(128, seed=31)).requires_grad_(True)
    beta = (0.1 * rnd(128, seed=32)).requires_grad_(True)
    dz = rnd(n, 128, seed=33)
    F.layer_norm(s, (128,), gamma, beta, 1e-5).backward(dz)
    sd = s.detach()
    mean = sd.mean(dim=1)
    rstd = torch.rsqrt(sd.var(dim=1, unbiased=False) + 1e-5)
    dg, db = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda")
    colsum = torch.zeros(128, device="cuda")
    ds = layernorm_bwd(dz, sd, torch.stack([mean, rstd], dim=1).contiguous(), gamma.detach(), dg, db, colsum)
    torch.testing.assert_close(ds, s.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(colsum, s.grad.sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(dg, gamma.grad, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, beta.grad, rtol=1e-4, atol=1e-3)


def test_bf16_operand_storage_is_equivalent():
    """precision 1 with bf16 A rows / bf16 output rows == the fp32-storage call on the same values (the kernel rounds its
    operands to bf16 anyway); the bf16 output is the rounded fp32 output."""
    from geomae_b200.dense import tc_linear, tc_wgrad
    from geomae_b200 import lib as L
    import ctypes as C
    n = 3001
    x = rnd(n, 256, seed=40)
    w = rnd(128, 256, seed=41) * 0.1
    ref = tc_linear(x.bfloat16().float(), w, n_out=128, precision=1)
    # bf16 A rows
    a = L.LinearArgs()
    x16 = x.bfloat16().contiguous()
    out = torch.empty(n, 128, device="cuda")
    a.A, a.lda, a.n_rows, a.K = x16.data_ptr(), 256, n, 256
    a.W, a.ldw, a.w_rows, a.w_mn_major = w.data_ptr(), 256, 128, 0
    a.N_total, a.out, a.ldo, a.precision, a.a_bf16 = 128, out.data_ptr(), 128, 1, 1
    L.run("tc_linear", C.byref(a), L.stream_ptr(x.device))
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    # bf16 output rows
    out16 = torch.empty(n, 128, device="cuda", dtype=torch.bfloat16)
    a.out, a.out_bf16 = out16.data_ptr(), 1
    L.run("tc_linear", C.byref(a), L.stream_ptr(x.device))
    torch.testing.assert_close(out16, ref.bfloat16(), rtol=0, atol=0)
    # bf16 dY rows in the weight gradient
    dy = rnd(n, 128, seed=42)
    xx = rnd(n, 128, seed=43)
    dw_ref = torch.zeros(128, 128, device="cuda")
    tc_wgrad(dy.bfloat16().float(), xx, dw_ref, None, precision=1)
    dw = torch.zeros(128, 128, device="cuda")
    g = L.WgradArgs()
    dy16 = dy.bfloat16().contiguous()
    g.dY, g.ldy, g.X, g.ldx, g.n_rows = dy16.data_ptr(), 128, xx.data_ptr(), 128, n
    g.dW, g.ldw, g.M_total, g.N_total, g.precision, g.dy_bf16 = dw.data_ptr(), 128, 128, 128, 1, 1
    L.run("tc_wgrad", C.byref(g), L.stream_ptr(x.device))
    torch.testing.assert_close(dw, dw_ref, rtol=1e-4, atol=1e-3)
