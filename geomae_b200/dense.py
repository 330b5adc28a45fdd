"""Host wrappers of the tensor-core dense kernels (csrc/sra_layer.cu) used by the SRA block."""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L


def pack_weight(W, want_lo=True):
    """bf16 tensor-core image(s) of a [rows, cols] fp32 weight (cols % 64 == 0): (hi, lo | None) uint8 tensors made of
    swizzled [128 x 64] blocks, ready for bulk copies into shared memory (geomae_pack_weights)."""
    rows, cols = W.shape
    assert W.is_contiguous() and cols % 64 == 0
    nbytes = (rows + 127) // 128 * 128 * cols * 2
    hi = torch.empty(nbytes, dtype=torch.uint8, device=W.device)
    lo = torch.empty(nbytes, dtype=torch.uint8, device=W.device) if want_lo else None
    Wp, r, c = (C.c_void_p * 1)(W.data_ptr()), (C.c_int32 * 1)(rows), (C.c_int32 * 1)(cols)
    hp = (C.c_void_p * 1)(hi.data_ptr())
    lp = (C.c_void_p * 1)(lo.data_ptr()) if want_lo else None
    L.run("pack_weights", 1, Wp, r, c, hp, lp, L.stream_ptr(W.device))
    return hi, lo


def tc_linear(A, W, *, n_out, w_mn_major=False, bias=None, pos_table=None, tok_cell=None, pos_slabs=0, a_gelu=False,
              add_src=None, ln=None, gelu_u=None, out=None, precision=3, packed=None):
    """out = prologue(A) @ (W^T | W) + epilogue on tcgen05 (see geomae_linear_args in include/geomae_b200.h).

    ln = (gamma, beta, eps, want_saved) selects the LayerNorm epilogue and returns (out, ln_in, ln_stats)."""
    L.require_cuda(A, "A")
    n, K = A.shape
    assert A.stride(1) == 1 and W.stride(1) == 1
    if out is None:
        out = torch.empty((n, n_out), dtype=torch.float32, device=A.device)
    a = L.LinearArgs()
    a.A, a.lda, a.n_rows, a.K = A.data_ptr(), A.stride(0), n, K
    a.pos_table = pos_table.data_ptr() if pos_table is not None else None
    a.tok_cell = tok_cell.data_ptr() if tok_cell is not None else None
    a.pos_slabs, a.a_gelu = pos_slabs, int(a_gelu)
    a.W, a.ldw, a.w_rows, a.w_mn_major = W.data_ptr(), W.stride(0), W.shape[0], int(w_mn_major)
    a.bias = bias.data_ptr() if bias is not None else None
    a.N_total, a.out, a.ldo = n_out, out.data_ptr(), out.stride(0)
    if add_src is not None:
        a.add_src, a.ld_add = add_src.data_ptr(), add_src.stride(0)
    ln_in = ln_stats = None
    a.epilogue = 0
    if ln is not None:
        gamma, beta, eps, want_saved = ln
        a.ln_gamma, a.ln_beta, a.ln_eps = gamma.data_ptr(), beta.data_ptr(), eps
        if want_saved:
            ln_in = torch.empty((n, n_out), dtype=torch.float32, device=A.device)
            ln_stats = torch.empty((n, 2), dtype=torch.float32, device=A.device)
            a.ln_in, a.ln_stats = ln_in.data_ptr(), ln_stats.data_ptr()
        a.epilogue = 1
    if gelu_u is not None:
        a.gelu_u, a.ldu, a.epilogue = gelu_u.data_ptr(), gelu_u.stride(0), 2
    a.precision = precision
    if packed is not None:
        a.Wp_hi = packed[0].data_ptr()
        a.Wp_lo = packed[1].data_ptr() if packed[1] is not None else None
    L.run("tc_linear", C.byref(a), L.stream_ptr(A.device))
    if ln is not None:
        return out, ln_in, ln_stats
    return out


def tc_wgrad(dY, X, dW, db=None, *, pos_table=None, tok_cell=None, pos_slabs=0, x_gelu=False, precision=3):
    """dW += dY^T @ prologue(X); db += dY.sum(0).  dW/db are fp32 accumulators (e.g. views of the flat grad buffer)."""
    n, m_total = dY.shape
    a = L.WgradArgs()
    a.dY, a.ldy, a.X, a.ldx, a.n_rows = dY.data_ptr(), dY.stride(0), X.data_ptr(), X.stride(0), n
    a.pos_table = pos_table.data_ptr() if pos_table is not None else None
    a.tok_cell = tok_cell.data_ptr() if tok_cell is not None else None
    a.pos_slabs, a.x_gelu = pos_slabs, int(x_gelu)
    a.dW, a.ldw = dW.data_ptr(), dW.stride(0)
    a.db = db.data_ptr() if db is not None else None
    a.M_total, a.N_total, a.precision = m_total, X.shape[1], precision
    assert tuple(dW.shape) == (m_total, X.shape[1])
    L.run("tc_wgrad", C.byref(a), L.stream_ptr(dY.device))


def layernorm_bwd(d_out, ln_in, ln_stats, gamma, d_gamma, d_beta, d_in_colsum=None):
    """-> d_in; d_gamma / d_beta (and, when given, the column sums of d_in) are accumulated in place."""
    d_in = torch.empty_like(d_out)
    L.run("layernorm_bwd", L.ptr(d_out), L.ptr(ln_in), L.ptr(ln_stats), L.ptr(gamma), d_out.shape[0], d_out.shape[1],
          L.ptr(d_in), L.ptr(d_gamma), L.ptr(d_beta), L.ptr(d_in_colsum), L.stream_ptr(d_out.device))
    return d_in


class _TCLinearFn(torch.autograd.Function):
    """y = x W^T (+ b) on tcgen05; backward = input gradient (MN-major weight view) + weight gradient accumulated
    straight into ``weight.grad`` (and ``bias.grad``)."""

    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        x = x.contiguous()
        y = tc_linear(x, weight, n_out=weight.shape[0], bias=bias, precision=precision)
        ctx.save_for_backward(x)
        ctx.weight, ctx.bias, ctx.precision = weight, bias, precision
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w, b = ctx.weight, ctx.bias
        dy = dy.contiguous()
        for p in (w, b):
            if p is not None and p.grad is None:
                p.grad = torch.zeros_like(p)
        tc_wgrad(dy, x, w.grad, b.grad if b is not None else None, precision=ctx.precision)
        dx = tc_linear(dy, w, n_out=w.shape[1], w_mn_major=True, precision=ctx.precision) if ctx.needs_input_grad[0] \
            else None
        return dx, None, None, None


def tc_linear_module(x, linear: torch.nn.Linear, precision):
    """nn.Linear forward/backward on the tensor cores (in/out features must be multiples of 128)."""
    return _TCLinearFn.apply(x, linear.weight, linear.bias, precision)


def sra_chain_fwd(x, *, attn=None, layer=None, next_in_proj=None, pos_table=None, tok_cell_next=None):
    """Fused forward chain of one EncoderLayer on 128-token tiles (geomae_sra_chain_fwd, csrc/sra_chain.cu).

    layer = dict(Wo, bo, W1, b1, W2, b2, g1, be1, g2, be2, eps) runs out-proj+LN1 -> FFN -> LN2 on `attn` (bf16 [n,128]);
    next_in_proj = (Win [384,128], bin [384]) appends the next layer's in-projection of (z + pos | z) (or of x itself
    when layer is None: the stack prologue).  Returns a dict of the produced tensors."""
    n = x.shape[0]
    dev = x.device
    a = L.ChainFwdArgs()
    a.n_tokens, a.mode, a.x = n, (1 if layer is not None else 0) | (2 if next_in_proj is not None else 0), x.data_ptr()
    out, keep = {}, []
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)       # noqa: E731
    b16 = lambda *s: torch.empty(s, dtype=torch.bfloat16, device=dev)      # noqa: E731
    if layer is not None:
        a.attn = attn.data_ptr()
        for key, field in (("Wo", "p_out_proj"), ("W1", "p_lin1"), ("W2", "p_lin2")):
            img = pack_weight(layer[key], want_lo=False)[0]
            keep.append(img)
            setattr(a, field, img.data_ptr())
        for key, field in (("bo", "out_proj_b"), ("b1", "lin1_b"), ("b2", "lin2_b"), ("g1", "norm1_w"), ("be1", "norm1_b"),
                           ("g2", "norm2_w"), ("be2", "norm2_b")):
            setattr(a, field, layer[key].data_ptr())
        a.ln_eps = layer["eps"]
        out.update(st1=f32(n, 2), st2=f32(n, 2), z=f32(n, 128), xh1_16=b16(n, 128), xh2_16=b16(n, 128), u16=b16(n, 256),
                   g16=b16(n, 256))
        for k in ("st1", "st2", "z", "xh1_16", "xh2_16", "u16", "g16"):
            setattr(a, k, out[k].data_ptr())
    else:
        out.update(xb16=b16(n, 128))
        a.xb16 = out["xb16"].data_ptr()
    if next_in_proj is not None:
        Win, bin_ = next_in_proj
        img = pack_weight(Win, want_lo=False)[0]
        keep.append(img)
        a.p_in_proj_next, a.in_proj_b_next = img.data_ptr(), bin_.data_ptr()
        a.pos_table, a.tok_cell_next = pos_table.data_ptr(), tok_cell_next.data_ptr()
        out.update(qkv16=b16(n, 384))
        a.qkv16_next = out["qkv16"].data_ptr()
    L.run("sra_chain_fwd", C.byref(a), L.stream_ptr(dev))
    return out
