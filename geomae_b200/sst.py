"""Sparse Regional Attention blocks — mirror of mmdet3d/models/sst/sst_basic_block.py with the
same module tree / state_dict keys (``win_attn.self_attn.in_proj_weight`` …), computing on CSR
windows instead of padded buckets.  The attention core is the hand-written kernel in
csrc/sra_attention.cu; nn.MultiheadAttention is kept purely as the parameter container."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import lib as L
from .windows import WindowLayout, pos_table


def _attn_fwd(qkv, win, n_heads, tc=False):
    """tc=False: fp32 SIMT kernel (parity mode); tc=True: bf16 tensor-core kernel (csrc/sra_attention_tc.cu; qkv may be
    fp32 or bf16 rows)."""
    n, three_d = qkv.shape
    out = torch.empty((n, three_d // 3), dtype=torch.float32, device=qkv.device)
    lse = torch.empty((n, n_heads), dtype=torch.float32, device=qkv.device)
    args = (L.ptr(qkv), n, n_heads, L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]), L.ptr(win["tok_win"]), L.ptr(out), L.ptr(lse))
    if tc:
        L.run("sra_attention_tc_fwd", *args, int(qkv.dtype == torch.bfloat16), L.stream_ptr(qkv.device))
    else:
        L.run("sra_attention_fwd", *args, L.stream_ptr(qkv.device))
    return out, lse


def _attn_bwd(qkv, out, lse, d_out, win, n_heads, tc=False, dd=None):
    d_qkv = torch.empty_like(qkv)
    if tc:
        flags = int(qkv.dtype == torch.bfloat16) | (2 if d_out.dtype == torch.bfloat16 else 0) | \
            (4 if d_qkv.dtype == torch.bfloat16 else 0)
        L.run("sra_attention_tc_bwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), L.ptr(d_out), qkv.shape[0], n_heads,
              L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]), L.ptr(win["tok_win"]), L.ptr(d_qkv), L.ptr(dd), flags,
              L.stream_ptr(qkv.device))
        return d_qkv
    scratch = torch.empty((qkv.shape[0], n_heads), dtype=torch.float32, device=qkv.device)
    L.run("sra_attention_bwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), L.ptr(d_out), qkv.shape[0], n_heads,
          L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]), L.ptr(win["tok_win"]), L.ptr(d_qkv), L.ptr(scratch),
          L.stream_ptr(qkv.device))
    return d_qkv


class _SRAAttention(torch.autograd.Function):
    """out[i] = softmax_j(q_i.k_j / sqrt(hd)) v_j over the tokens j sharing i's window."""

    @staticmethod
    def forward(ctx, qkv, win, n_heads, tc=False):
        qkv = qkv.contiguous()
        out, lse = _attn_fwd(qkv, win, n_heads, tc)
        ctx.save_for_backward(qkv, out, lse)
        ctx.win, ctx.n_heads, ctx.tc = win, n_heads, tc
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        return _attn_bwd(qkv, out, lse, d_out.contiguous(), ctx.win, ctx.n_heads, ctx.tc), None, None, None


def _grad_of(p):
    """The accumulator the fused backward adds into (FlatTrainer pre-binds it to the flat gradient buffer)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


class _SRALayerFn(torch.autograd.Function):
    """One whole EncoderLayer (sst_basic_block.py:85-102) as 5 forward / 11 backward kernels:
    in-proj(+pos) -> window attention -> out-proj+residual+LN1 -> FFN1 -> GELU+FFN2+residual+LN2, every GEMM on
    tcgen05.  Parameter gradients are accumulated straight into ``p.grad`` (no autograd bookkeeping for the 12
    parameter tensors); only d(input) is returned to autograd."""

    @staticmethod
    def forward(ctx, x, layer, win, table, precision):
        from .dense import tc_linear
        x = x.contiguous()
        mha = layer.win_attn.self_attn
        d, nh = layer.win_attn.d_model, layer.win_attn.nhead
        cell = win["tok_cell"]
        qkv = tc_linear(x, mha.in_proj_weight, n_out=3 * d, bias=mha.in_proj_bias, pos_table=table, tok_cell=cell,
                        pos_slabs=2, precision=precision)
        a, lse = _attn_fwd(qkv, win, nh, precision == 1)
        y, s1, st1 = tc_linear(a, mha.out_proj.weight, n_out=d, bias=mha.out_proj.bias, add_src=x,
                               ln=(layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, True), precision=precision)
        u = tc_linear(y, layer.linear1.weight, n_out=layer.linear1.out_features, bias=layer.linear1.bias,
                      precision=precision)
        z, s2, st2 = tc_linear(u, layer.linear2.weight, n_out=d, bias=layer.linear2.bias, add_src=y, a_gelu=True,
                               ln=(layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, True), precision=precision)
        ctx.save_for_backward(x, qkv, a, lse, s1, st1, y, u, s2, st2)
        ctx.layer, ctx.win, ctx.table, ctx.precision = layer, win, table, precision
        return z

    @staticmethod
    def backward(ctx, dz):
        from .dense import layernorm_bwd, tc_linear, tc_wgrad
        x, qkv, a, lse, s1, st1, y, u, s2, st2 = ctx.saved_tensors
        layer, win, table, prec = ctx.layer, ctx.win, ctx.table, ctx.precision
        mha = layer.win_attn.self_attn
        d, nh = layer.win_attn.d_model, layer.win_attn.nhead
        g = _grad_of
        dz = dz.contiguous()
        ds2 = layernorm_bwd(dz, s2, st2, layer.norm2.weight, g(layer.norm2.weight), g(layer.norm2.bias),
                            g(layer.linear2.bias))
        du = tc_linear(ds2, layer.linear2.weight, n_out=u.shape[1], w_mn_major=True, gelu_u=u, precision=prec)
        tc_wgrad(ds2, u, g(layer.linear2.weight), None, x_gelu=True, precision=prec)
        dy = tc_linear(du, layer.linear1.weight, n_out=d, w_mn_major=True, add_src=ds2, precision=prec)
        tc_wgrad(du, y, g(layer.linear1.weight), g(layer.linear1.bias), precision=prec)
        ds1 = layernorm_bwd(dy, s1, st1, layer.norm1.weight, g(layer.norm1.weight), g(layer.norm1.bias),
                            g(mha.out_proj.bias))
        da = tc_linear(ds1, mha.out_proj.weight, n_out=d, w_mn_major=True, precision=prec)
        tc_wgrad(ds1, a, g(mha.out_proj.weight), None, precision=prec)
        dqkv = _attn_bwd(qkv, a, lse, da, win, nh, prec == 1)
        dx = tc_linear(dqkv, mha.in_proj_weight, n_out=d, w_mn_major=True, add_src=ds1, precision=prec)
        tc_wgrad(dqkv, x, g(mha.in_proj_weight), g(mha.in_proj_bias), pos_table=table, tok_cell=win["tok_cell"],
                 pos_slabs=2, precision=prec)
        return dx, None, None, None, None


def sra_attention(qkv, win, n_heads, tc=False):
    L.require_cuda(qkv, "qkv")
    if qkv.dtype != torch.float32 and not (tc and qkv.dtype == torch.bfloat16):
        raise RuntimeError("sra_attention expects float32 q|k|v rows (bfloat16 only with tc=True)")
    return _SRAAttention.apply(qkv, win, n_heads, tc)


class WindowAttention(nn.Module):
    """sst_basic_block.py:13-61.  q = k = x + pos, v = x; heads of 16 channels."""

    def __init__(self, d_model, nhead, dropout, batch_first=False, layer_id=None):
        super().__init__()
        if dropout != 0:
            raise NotImplementedError("attention dropout is 0 on the GeoMAE path")
        if d_model != 16 * nhead:
            raise NotImplementedError("the SRA kernel is specialised for head_dim 16")
        self.nhead, self.d_model, self.layer_id = nhead, d_model, layer_id
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout)  # parameter container only

    def forward(self, feat_2d, pos, win):
        d = self.d_model
        w, b = self.self_attn.in_proj_weight, self.self_attn.in_proj_bias
        qk = F.linear(feat_2d + pos, w[:2 * d], b[:2 * d])
        v = F.linear(feat_2d, w[2 * d:], b[2 * d:])
        attn = sra_attention(torch.cat([qk, v], dim=1), win, self.nhead)
        return F.linear(attn, self.self_attn.out_proj.weight, self.self_attn.out_proj.bias)


class EncoderLayer(nn.Module):
    """sst_basic_block.py:63-102 — post-norm: x = LN(x + SRA(x)); x = LN(x + W2 gelu(W1 x))."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", batch_first=False,
                 layer_id=None, mlp_dropout=0):
        super().__init__()
        assert not batch_first
        if mlp_dropout != 0:
            raise NotImplementedError("mlp_dropout is 0 on the GeoMAE path")
        self.win_attn = WindowAttention(d_model, nhead, dropout, layer_id=layer_id)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.activation = {"relu": F.relu, "gelu": F.gelu}[activation]
        self.activation_name = activation
        # "tc3": fused tcgen05 kernels, bf16x3 split (fp32-parity);  "tc1": same kernels, plain bf16;
        # "glue": hand-written attention + library GEMM/LayerNorm under autograd (the round-1 v1 path)
        self.impl = "tc3"

    def forward(self, src, pos, win, table=None):
        if self.impl in ("tc3", "tc1"):
            if self.activation_name != "gelu":
                raise NotImplementedError("the fused SRA layer implements the GELU feed-forward of the GeoMAE configs")
            return _SRALayerFn.apply(src, self, win, table, 3 if self.impl == "tc3" else 1)
        src = self.norm1(src + self.win_attn(src, pos, win))
        src2 = self.linear2(self.activation(self.linear1(src)))
        return self.norm2(src + src2)


class BasicShiftBlock(nn.Module):
    """sst_basic_block.py:104-147 — layer 0 on shift-0 windows, layer 1 on shift-1 windows."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", batch_first=False,
                 block_id=-100):
        super().__init__()
        self.encoder_list = nn.ModuleList([
            EncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, batch_first, layer_id=block_id * 2 + j)
            for j in range(2)])

    def forward(self, src, layout: WindowLayout, pos_list, table=None):
        n_shifts = layout.spec.n_shifts
        for i, layer in enumerate(self.encoder_list):
            s = i % n_shifts
            src = layer(src, pos_list[s] if pos_list is not None else None, layout.shift(s), table)
        return src


class SRAStack:
    """A sequence of EncoderLayers executed by ONE C-ABI call per direction
    (geomae_sra_stack_forward / _backward, csrc/sra_stack.cu)."""

    def __init__(self, layers, shifts):
        self.layers, self.shifts = list(layers), list(shifts)
        self._packed = None     # bf16 weight images (hi, lo) per layer, refilled by the executor every forward

    def _pack_arena(self, device):
        if self._packed is None or self._packed.device != device:
            first = self.layers[0]
            d, f = first.win_attn.d_model, first.linear1.out_features
            per_layer = 2 * 2 * (3 * d * d + d * d + f * d + d * f)       # bytes: (hi, lo) x bf16
            self._per_layer = per_layer
            self._packed = torch.empty(per_layer * len(self.layers), dtype=torch.uint8, device=device)
        return self._packed

    def _fingerprint(self):
        """Cheap identity of everything `_structs` captures by pointer: the parameters and gradient accumulators of the
        first and last layer (FlatTrainer keeps all of them in two flat buffers, so they move together) + the arena."""
        a, b = self.layers[0].norm1.weight, self.layers[-1].linear2.weight
        return (a.data_ptr(), 0 if a.grad is None else a.grad.data_ptr(), b.data_ptr(),
                0 if b.grad is None else b.grad.data_ptr(), 0 if self._packed is None else self._packed.data_ptr())

    def _structs(self):
        cached = self.__dict__.get("_struct_cache")
        if cached is not None and cached[0] == self._fingerprint():
            return cached[1]
        arr = self._build_structs()
        self._struct_cache = (self._fingerprint(), arr)
        return arr

    def _build_structs(self):
        import ctypes as C
        arr = (L.SRALayer * len(self.layers))()
        for s, layer, shift in zip(arr, self.layers, self.shifts):
            mha = layer.win_attn.self_attn
            ps = [mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, layer.linear1.weight,
                  layer.linear1.bias, layer.linear2.weight, layer.linear2.bias, layer.norm1.weight, layer.norm1.bias,
                  layer.norm2.weight, layer.norm2.bias]
            s.shift, s.ln_eps = shift, layer.norm1.eps
            for name, gname, p in zip(L._LAYER_PARAMS, L._LAYER_GRADS, ps):
                setattr(s, name, p.data_ptr())
                setattr(s, gname, _grad_of(p).data_ptr())
        arena = self._pack_arena(self.layers[0].norm1.weight.device)
        d, f = self.layers[0].win_attn.d_model, self.layers[0].linear1.out_features
        sizes = [("p_in_proj", 3 * d * d), ("p_out_proj", d * d), ("p_lin1", f * d), ("p_lin2", d * f)]
        for i, s in enumerate(arr):
            off = arena.data_ptr() + i * self._per_layer
            for name, elems in sizes:
                getattr(s, name)[0] = off
                getattr(s, name)[1] = off + 2 * elems
                off += 4 * elems
        return arr

    def ctx(self, layout: WindowLayout, table, n, precision):
        first = self.layers[0]
        c = L.SRACtx()
        c.n_tokens, c.d_model, c.n_heads = n, first.win_attn.d_model, first.win_attn.nhead
        c.ffn, c.precision, c.pos_table = first.linear1.out_features, precision, table.data_ptr()
        for s in range(layout.spec.n_shifts):
            w = layout.shift(s)
            c.shift[s].win_ptr, c.shift[s].win_tok = w["win_ptr"].data_ptr(), w["win_tok"].data_ptr()
            c.shift[s].tok_win, c.shift[s].tok_cell = w["tok_win"].data_ptr(), w["tok_cell"].data_ptr()
        if precision == 1:      # gathered bf16 position rows per shift (filled by the forward executor), kept on the layout
            pos16 = layout.__dict__.setdefault("_pos16", {})
            key = (n, table.data_ptr())
            if key not in pos16:
                pos16[key] = torch.empty((layout.spec.n_shifts, max(n, 1), c.d_model), dtype=torch.bfloat16, device=table.device)
            for s in range(layout.spec.n_shifts):
                c.pos16[s] = pos16[key][s].data_ptr()
        return c

    def __call__(self, x, layout, table, precision):
        return _SRAStackFn.apply(x, self, layout, table, precision)


def _carve(n, d, f, heads, n_layers, device):
    """One arena for everything a stack saves for backward; returns (arena, SRASaved array, last z view)."""
    sizes = [("qkv", 3 * d), ("attn", d), ("lse", heads), ("s1", d), ("st1", 2), ("y", d), ("u", f), ("s2", d),
             ("st2", 2), ("z", d), ("g", f // 2), ("xp", d // 2), ("xb", d // 2)]     # last three: bf16 rows (bf16 mode)
    pad = lambda k: (k + 63) // 64 * 64  # noqa: E731
    per_layer = sum(pad(n * w) for _, w in sizes)
    arena = torch.empty(max(per_layer * n_layers, 1), dtype=torch.float32, device=device)
    saved = (L.SRASaved * n_layers)()
    base = arena.data_ptr()
    z_last = None
    for l in range(n_layers):
        off = l * per_layer
        for name, w in sizes:
            setattr(saved[l], name, base + 4 * off)
            if name == "z" and l == n_layers - 1:
                z_last = arena[off:off + n * d].view(n, d)
            off += pad(n * w)
    return arena, saved, z_last


class _SRAStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stack: SRAStack, layout, table, precision):
        import ctypes as C
        x = x.contiguous()
        n = x.shape[0]
        first = stack.layers[0]
        d, f, nh = first.win_attn.d_model, first.linear1.out_features, first.win_attn.nhead
        arena, saved, z = _carve(n, d, f, nh, len(stack.layers), x.device)
        layers = stack._structs()
        c = stack.ctx(layout, table, n, precision)
        L.run("sra_stack_forward", C.byref(c), len(stack.layers), layers, saved, L.ptr(x), L.stream_ptr(x.device))
        L.add_launches((2 * len(stack.layers) + 1 if precision == 1 else 5 * len(stack.layers)) - 1 + 1)   # + weight packing
        ctx.save_for_backward(x, arena)
        ctx.stack, ctx.saved, ctx.layers, ctx.c, ctx.layout = stack, saved, layers, c, layout
        return z

    @staticmethod
    def backward(ctx, dz):
        import ctypes as C
        x, arena = ctx.saved_tensors
        n, d = x.shape
        f = ctx.c.ffn
        dz = dz.contiguous()
        dx = torch.empty_like(x)
        scratch = torch.empty(max(L.lib().geomae_sra_scratch_floats(C.byref(ctx.c), 1), 1), dtype=torch.float32,
                              device=x.device)
        L.run("sra_stack_backward", C.byref(ctx.c), len(ctx.stack.layers), ctx.layers, ctx.saved, L.ptr(x), L.ptr(dz),
              L.ptr(dx), L.ptr(scratch), L.stream_ptr(x.device))
        nl = len(ctx.stack.layers)
        L.add_launches(3 * nl if ctx.c.precision == 1 else 10 * nl)    # chain + attention + wgrad (+ final dx: the call itself)
        return dx, None, None, None, None


class _SRADualStackFn(torch.autograd.Function):
    """Two stacks over the same input (centroid / density decoders) executed concurrently."""

    @staticmethod
    def forward(ctx, x, stack_a: SRAStack, stack_b: SRAStack, layout, table, precision):
        import ctypes as C
        x = x.contiguous()
        n = x.shape[0]
        first = stack_a.layers[0]
        nl = len(stack_a.layers)
        assert nl == len(stack_b.layers) and stack_a.shifts == stack_b.shifts
        d, f, nh = first.win_attn.d_model, first.linear1.out_features, first.win_attn.nhead
        arena_a, saved_a, za = _carve(n, d, f, nh, nl, x.device)
        arena_b, saved_b, zb = _carve(n, d, f, nh, nl, x.device)
        la, lb = stack_a._structs(), stack_b._structs()
        c = stack_a.ctx(layout, table, n, precision)
        L.run("sra_stack2_forward", C.byref(c), nl, la, saved_a, lb, saved_b, L.ptr(x), L.stream_ptr(x.device))
        L.add_launches((2 * (2 * nl + 2) if precision == 1 else 2 * (5 * nl + 1)) - 1)
        ctx.save_for_backward(x, arena_a, arena_b)
        ctx.keep = (saved_a, saved_b, la, lb, c, layout, nl)
        return za, zb

    @staticmethod
    def backward(ctx, dza, dzb):
        import ctypes as C
        x, _, _ = ctx.saved_tensors
        saved_a, saved_b, la, lb, c, _, nl = ctx.keep
        dza, dzb = dza.contiguous(), dzb.contiguous()
        dxa, dxb = torch.empty_like(x), torch.empty_like(x)
        scratch = torch.empty(max(L.lib().geomae_sra_scratch_floats(C.byref(c), 2), 1), dtype=torch.float32,
                              device=x.device)
        L.run("sra_stack2_backward", C.byref(c), nl, la, saved_a, lb, saved_b, L.ptr(x), L.ptr(dza), L.ptr(dzb), L.ptr(dxa),
              L.ptr(dxb), L.ptr(scratch), L.stream_ptr(x.device))
        L.add_launches((2 * (3 * nl + 1) if c.precision == 1 else 20 * nl + 2) - 1)
        return dxa + dxb, None, None, None, None, None


def window_pos_embed(layout: WindowLayout, d_model, temperature):
    """Per-token position rows for each shift: table[tok_cell] (…top_only.py:361-399)."""
    table = pos_table(layout.spec.window_shape, d_model, temperature, layout.tok_cell.device)
    return [table.index_select(0, layout.tok_cell[s, :layout.n_tokens].long()) for s in range(layout.spec.n_shifts)]
