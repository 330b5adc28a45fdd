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


class _SRAAttention(torch.autograd.Function):
    """out[i] = softmax_j(q_i.k_j / sqrt(hd)) v_j over the tokens j sharing i's window."""

    @staticmethod
    def forward(ctx, qkv, win, n_heads):
        qkv = qkv.contiguous()
        n, three_d = qkv.shape
        d = three_d // 3
        out = torch.empty((n, d), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((n, n_heads), dtype=torch.float32, device=qkv.device)
        L.run("sra_attention_fwd", L.ptr(qkv), n, n_heads, L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]),
                                                 L.ptr(win["n_windows"]), win["max_windows"], L.ptr(out), L.ptr(lse),
                                                 L.stream_ptr(qkv.device))
        ctx.save_for_backward(qkv, out, lse)
        ctx.win, ctx.n_heads = win, n_heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        d_out = d_out.contiguous()
        d_qkv = torch.empty_like(qkv)
        win = ctx.win
        L.run("sra_attention_bwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), L.ptr(d_out), qkv.shape[0],
                                                 ctx.n_heads, L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]),
                                                 L.ptr(win["n_windows"]), win["max_windows"], L.ptr(d_qkv),
                                                 L.stream_ptr(qkv.device))
        return d_qkv, None, None


def sra_attention(qkv, win, n_heads):
    L.require_cuda(qkv, "qkv")
    if qkv.dtype != torch.float32:
        raise RuntimeError("sra_attention expects float32 q|k|v rows")
    return _SRAAttention.apply(qkv, win, n_heads)


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

    def forward(self, src, pos, win):
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

    def forward(self, src, layout: WindowLayout, pos_list):
        n_shifts = layout.spec.n_shifts
        for i, layer in enumerate(self.encoder_list):
            s = i % n_shifts
            src = layer(src, pos_list[s], layout.shift(s))
        return src


def window_pos_embed(layout: WindowLayout, d_model, temperature):
    """Per-token position rows for each shift: table[tok_cell] (…top_only.py:361-399)."""
    table = pos_table(layout.spec.window_shape, d_model, temperature, layout.tok_cell.device)
    return [table.index_select(0, layout.tok_cell[s, :layout.n_tokens].long()) for s in range(layout.spec.n_shifts)]
