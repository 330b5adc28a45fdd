"""Fused token-local SRA chain kernels (csrc/sra_chain.cu) against torch fp32 with the kernel's rounding points
(bf16 operands of every tensor-core product, fp32 accumulation / LayerNorm / GELU / residuals)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def bf(t):
    return t.to(torch.bfloat16).float()


def make_layer(seed):
    return dict(Wo=rnd(128, 128, scale=0.1, seed=seed), bo=rnd(128, scale=0.1, seed=seed + 1),
                W1=rnd(256, 128, scale=0.1, seed=seed + 2), b1=rnd(256, scale=0.1, seed=seed + 3),
                W2=rnd(128, 256, scale=0.1, seed=seed + 4), b2=rnd(128, scale=0.1, seed=seed + 5),
                g1=1 + 0.1 * rnd(128, seed=seed + 6), be1=0.1 * rnd(128, seed=seed + 7),
                g2=1 + 0.1 * rnd(128, seed=seed + 8), be2=0.1 * rnd(128, seed=seed + 9), eps=1e-5)


def reference(x, attn, lay, nxt, table, cell):
    out = {}
    z = x
    if lay is not None:
        s1 = x + attn.float() @ bf(lay["Wo"]).T + lay["bo"]
        y = F.layer_norm(s1, (128,), lay["g1"], lay["be1"], lay["eps"])
        u = bf(y) @ bf(lay["W1"]).T + lay["b1"]
        g = F.gelu(u)
        s2 = y + bf(g) @ bf(lay["W2"]).T + lay["b2"]
        z = F.layer_norm(s2, (128,), lay["g2"], lay["be2"], lay["eps"])
        xh = lambda t: (t - t.mean(1, keepdim=True)) * torch.rsqrt(t.var(1, unbiased=False, keepdim=True) + lay["eps"])   # noqa: E731
        out.update(s1=s1, y=y, u=u, g=g, s2=s2, z=z, xh1=xh(s1), xh2=xh(s2),
                   st1=torch.stack([s1.mean(1), torch.rsqrt(s1.var(1, unbiased=False) + lay["eps"])], 1),
                   st2=torch.stack([s2.mean(1), torch.rsqrt(s2.var(1, unbiased=False) + lay["eps"])], 1))
    if nxt is not None:
        Win, bin_ = nxt
        xp, xb = bf(z + table[cell.long()]), bf(z)
        out.update(xp=xp, xb=xb, qkv=torch.cat([xp @ bf(Win[:256]).T + bin_[:256], xb @ bf(Win[256:]).T + bin_[256:]], 1))
    return out


def close(got, ref, what, atol, mean_tol):
    """max error within atol + 2 bf16 ulps of the value, mean error within mean_tol + half a bf16 ulp of the mean magnitude."""
    d = (got.float() - ref).abs()
    assert torch.isfinite(got.float()).all(), what
    excess = (d - 2.0 ** -7 * ref.abs()).max().item()
    assert excess <= atol, (what, "max", excess, d.max().item())
    assert d.mean().item() <= mean_tol + 2.0 ** -9 * ref.abs().mean().item(), (what, "mean", d.mean().item())


# 100: one partial tile; 1000: several tiles; 148*128+37: more tiles than CTAs (persistent loop, phase flips)
@pytest.mark.parametrize("n", [100, 1000, 148 * 128 * 2 + 37])
@pytest.mark.parametrize("mode", [3, 1, 2])
def test_chain_forward(n, mode):
    from geomae_b200.dense import sra_chain_fwd
    x = rnd(n, 128, seed=1)
    attn = rnd(n, 128, seed=2).to(torch.bfloat16)
    lay = make_layer(10) if mode & 1 else None
    nxt = (rnd(384, 128, scale=0.1, seed=30), rnd(384, scale=0.1, seed=31)) if mode & 2 else None
    table = rnd(144, 128, seed=4)
    cell = torch.randint(0, 144, (n,), dtype=torch.int32, device="cuda")
    got = sra_chain_fwd(x, attn=attn, layer=lay, next_in_proj=nxt, pos_table=table, tok_cell_next=cell)
    torch.cuda.synchronize()
    ref = reference(x, attn, lay, nxt, table, cell)
    if lay is not None:
        close(got["st1"], ref["st1"], "st1", 2e-4, 2e-5)
        close(got["xh1_16"], ref["xh1"], "xh1_16", 4e-2, 4e-3)  # bf16 storage: half an ulp of |xhat| <= 4
        # downstream of a bf16 rounding of y / g a last-bit difference in fp32 flips single operand roundings
        close(got["u16"], ref["u"], "u16", 4e-2, 3e-3)
        close(got["g16"], ref["g"], "g16", 4e-2, 3e-3)
        close(got["xh2_16"], ref["xh2"], "xh2_16", 4e-2, 4e-3)
        close(got["z"], ref["z"], "z", 2e-2, 2e-4)
        close(got["st2"][:, 0], ref["st2"][:, 0], "st2 mean", 2e-3, 2e-5)
    else:
        close(got["xb16"], ref["xb"], "xb16", 6e-2, 3e-4)
    if nxt is not None:
        close(got["qkv16"], ref["qkv"], "qkv16", 8e-2, 6e-3)


@pytest.mark.parametrize("n", [100, 1000, 148 * 128 + 37])
@pytest.mark.parametrize("first_layer", [False, True])
def test_wgrad_layer_tma(n, first_layer):
    """All weight / bias gradients of a layer from bf16 operands in one TMA-fed launch; accumulates into the buffers.
    The LayerNorm outputs are applied in the flush: y = xh1 * g1 + b1, x = xin * scale + shift (or x = xin)."""
    import ctypes as C
    from geomae_b200 import lib as L
    b16 = lambda cols, seed: rnd(n, cols, seed=seed).to(torch.bfloat16)      # noqa: E731
    t = dict(ds2_16=b16(128, 1), g16=b16(256, 2), du16=b16(256, 3), xh1_16=b16(128, 4), ds1_16=b16(128, 5),
             attn16=b16(128, 6), dqkv16=b16(384, 7), xin16=b16(128, 8), pos16=b16(128, 9))
    g1, b1 = 1 + 0.2 * rnd(128, seed=20), 0.3 * rnd(128, seed=21)
    sc, sh = (None, None) if first_layer else (1 + 0.2 * rnd(128, seed=22), 0.3 * rnd(128, seed=23))
    init = 0.5
    g = dict(g_lin2_w=torch.full((128, 256), init, device="cuda"), g_lin1_w=torch.full((256, 128), init, device="cuda"),
             g_lin1_b=torch.full((256,), init, device="cuda"), g_out_proj_w=torch.full((128, 128), init, device="cuda"),
             g_in_proj_w=torch.full((384, 128), init, device="cuda"), g_in_proj_b=torch.full((384,), init, device="cuda"),
             g_lin2_b=torch.full((128,), init, device="cuda"), g_out_proj_b=torch.full((128,), init, device="cuda"))
    a = L.WgradLayerArgs()
    a.n_tokens = n
    for k, v in {**t, **g}.items():
        setattr(a, k, v.data_ptr())
    a.norm1_w, a.norm1_b = g1.data_ptr(), b1.data_ptr()
    if not first_layer:
        a.in_scale, a.in_shift = sc.data_ptr(), sh.data_ptr()
    L.run("sra_wgrad_layer", C.byref(a), L.stream_ptr(torch.device("cuda")))
    torch.cuda.synchronize()
    f = {k: v.double() for k, v in t.items()}
    y = f["xh1_16"] * g1.double() + b1.double()
    x = f["xin16"] if first_layer else f["xin16"] * sc.double() + sh.double()
    xp = x + f["pos16"]
    ref = dict(g_lin2_w=f["ds2_16"].T @ f["g16"], g_lin1_w=f["du16"].T @ y, g_lin1_b=f["du16"].sum(0),
               g_out_proj_w=f["ds1_16"].T @ f["attn16"],
               g_in_proj_w=torch.cat([f["dqkv16"][:, :256].T @ xp, f["dqkv16"][:, 256:].T @ x], 0),
               g_in_proj_b=f["dqkv16"].sum(0), g_lin2_b=f["ds2_16"].sum(0), g_out_proj_b=f["ds1_16"].sum(0))
    scale = (n ** 0.5)
    for k in g:
        d = (g[k].double() - init - ref[k]).abs().max().item()
        assert d <= 4e-5 * scale + 2e-4, (k, d)


def test_pos_rows_bf16():
    from geomae_b200 import lib as L
    n = 1000
    table = rnd(144, 128, seed=1)
    cell = torch.randint(0, 144, (n,), dtype=torch.int32, device="cuda")
    out = torch.empty(n, 128, dtype=torch.bfloat16, device="cuda")
    L.run("pos_rows_bf16", L.ptr(table), L.ptr(cell), n, L.ptr(out), L.stream_ptr(torch.device("cuda")))
    torch.cuda.synchronize()
    assert torch.equal(out, table[cell.long()].to(torch.bfloat16))


def ln_bwd_ref(dz, s, gamma, eps):
    """LayerNorm backward as the kernel evaluates it: xhat rounded to bf16 (the saved tile), everything else fp32."""
    mean = s.mean(1, keepdim=True)
    rstd = torch.rsqrt(s.var(1, unbiased=False, keepdim=True) + eps)
    xh = bf((s - mean) * rstd)
    g = dz * gamma
    ds = rstd * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True))
    return ds, xh, torch.cat([mean, rstd], 1)


@pytest.mark.parametrize("n", [100, 1000, 148 * 128 * 2 + 37])
@pytest.mark.parametrize("mode", [3, 2, 1])
def test_chain_backward(n, mode):
    import ctypes as C
    from geomae_b200 import lib as L
    from geomae_b200.dense import pack_weight
    up, chain = bool(mode & 1), bool(mode & 2)
    lay = make_layer(10)
    Win = rnd(384, 128, scale=0.1, seed=30)
    dqkv = rnd(n, 384, seed=40).to(torch.bfloat16)
    ds1_up, dz_in = rnd(n, 128, seed=41), rnd(n, 128, seed=42)
    s2, s1 = rnd(n, 128, seed=43) + 0.3, rnd(n, 128, seed=44) - 0.2
    u16 = rnd(n, 256, seed=45).to(torch.bfloat16)
    attn16 = rnd(n, 128, seed=46).to(torch.bfloat16)
    dev = torch.device("cuda")
    f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)     # noqa: E731
    b16 = lambda *s: torch.zeros(s, dtype=torch.bfloat16, device=dev)    # noqa: E731
    # ---- reference
    dz = dqkv.float() @ bf(Win) + ds1_up if up else dz_in
    if chain:
        ds2, xh2, st2 = ln_bwd_ref(dz, s2, lay["g2"], lay["eps"])
        uu = u16.float().double()
        gelu_grad = (0.5 * (1 + torch.erf(uu / 2 ** 0.5)) + uu * torch.exp(-uu * uu / 2) / (2 * torch.pi) ** 0.5).float()
        du = (bf(ds2) @ bf(lay["W2"])) * gelu_grad
        dy = bf(du) @ bf(lay["W1"]) + ds2
        ds1, xh1, st1 = ln_bwd_ref(dy, s1, lay["g1"], lay["eps"])
        dO = bf(ds1) @ bf(lay["Wo"])
        dd = (dO * attn16.float()).view(n, 8, 16).sum(-1)
    # ---- kernel
    a = L.ChainBwdArgs()
    a.n_tokens, a.mode = n, mode
    keep = []
    if up:
        img = pack_weight(Win, want_lo=False)[0]
        keep.append(img)
        a.dqkv16_up, a.ds1_up, a.p_in_proj_up = dqkv.data_ptr(), ds1_up.data_ptr(), img.data_ptr()
    else:
        a.dz_in = dz_in.data_ptr()
    out = {}
    if chain:
        for key, field in (("W2", "p_lin2"), ("W1", "p_lin1"), ("Wo", "p_out_proj")):
            img = pack_weight(lay[key], want_lo=False)[0]
            keep.append(img)
            setattr(a, field, img.data_ptr())
        st2c, st1c = st2.contiguous(), st1.contiguous()
        xh2_16, xh1_16 = xh2.to(torch.bfloat16).contiguous(), xh1.to(torch.bfloat16).contiguous()
        keep += [st2c, st1c, xh2_16, xh1_16]
        a.xh2_16, a.st2, a.xh1_16, a.st1 = xh2_16.data_ptr(), st2c.data_ptr(), xh1_16.data_ptr(), st1c.data_ptr()
        a.u16, a.attn16, a.norm2_w, a.norm1_w = u16.data_ptr(), attn16.data_ptr(), lay["g2"].data_ptr(), lay["g1"].data_ptr()
        out = dict(ds2_16=b16(n, 128), du16=b16(n, 256), ds1_16=b16(n, 128), dattn16=b16(n, 128), ds1=f32(n, 128),
                   dd=f32(n, 8), g_norm2_w=f32(128), g_norm2_b=f32(128), g_norm1_w=f32(128), g_norm1_b=f32(128))
        for k, v in out.items():
            setattr(a, k, v.data_ptr())
    else:
        out = dict(dx=f32(n, 128))
        a.dx = out["dx"].data_ptr()
    L.run("sra_chain_bwd", C.byref(a), L.stream_ptr(dev))
    torch.cuda.synchronize()
    if not chain:
        close(out["dx"], dz, "dx", 2e-3, 2e-5)
        return
    close(out["ds2_16"], ds2, "ds2_16", 6e-2, 3e-3)
    close(out["du16"], du, "du16", 6e-2, 3e-3)
    close(out["ds1"], ds1, "ds1", 3e-2, 4e-4)
    close(out["ds1_16"], ds1, "ds1_16", 8e-2, 4e-3)
    close(out["dattn16"], dO, "dattn16", 6e-2, 3e-3)
    close(out["dd"], dd, "dd", 0.3, 3e-3)          # 16-term dots of O(5) gradients: rare operand-rounding flips upstream
    s = n ** 0.5
    for key, ref in (("g_norm2_w", (dz * xh2).sum(0)), ("g_norm2_b", dz.sum(0)), ("g_norm1_w", (dy * xh1).sum(0)),
                     ("g_norm1_b", dy.sum(0))):
        d = (out[key] - ref).abs().max().item()
        assert d <= 5e-3 * s + 1e-3, (key, d)
