"""MultiMAESSTSPChoose — mirror of mmdet3d/models/backbones/multi_mae_sst_spearate_top_only.py
(same registry key, ctor kwargs and state_dict keys).  Window bookkeeping is the CSR layout of
``windows.py``; there is no padding, no drop (a 12x12 window never exceeds the largest bucket,
SURVEY F9) and no host synchronisation inside the blocks."""
from __future__ import annotations

import torch
from torch import nn

from .registry import BACKBONES
from .sst import BasicShiftBlock, SRAStack, window_pos_embed
from .windows import pos_table
from .voxel import PillarBatch, VoxelGeometry
from .windows import WindowLayout, WindowSpec


class FusedHeads:
    """The six prediction heads (…top_only.py:279-300) as TWO tensor-core GEMMs: the five heads that read the
    centroid decoder share one [768, 128] weight (rows: low | cls_low | med | cls_med | top | zero pad), the normal
    head of the density decoder is padded to 128 rows.  Outputs stay column slices of the fused result — the loss
    kernel reads them with a row stride — and the backward is one dX GEMM + one weight-gradient GEMM per decoder."""

    def __init__(self, bb):
        self.heads_c = [bb.decoder_pred_low, bb.cls_pred_low, bb.decoder_pred_med, bb.cls_pred_med, bb.decoder_pred_top]
        self.head_d = bb.decoder_pred_density_top
        self.offsets, off = [], 0
        for h in self.heads_c:
            self.offsets.append(off)
            off += h.out_features
        self.n_c = off                                  # 723 for the GeoMAE config
        self.width_c = (off + 127) // 128 * 128         # 768
        self.width_d = 128

    def cat_params(self, dev):
        """The fused weights / biases of this step.  The buffers persist (pad rows are zeroed once, at creation); only
        the live rows are refreshed from the parameters."""
        buf = self.__dict__.get("_param_buf")
        if buf is None or buf[0].device != dev:
            z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)     # noqa: E731
            buf = self._param_buf = (z(self.width_c, 128), z(self.width_c), z(self.width_d, 128), z(self.width_d))
        wc, bc, wd, bd = buf
        torch.cat([h.weight.detach() for h in self.heads_c], out=wc[:self.n_c])
        torch.cat([h.bias.detach() for h in self.heads_c], out=bc[:self.n_c])
        wd[:self.head_d.out_features].copy_(self.head_d.weight.detach())
        bd[:self.head_d.out_features].copy_(self.head_d.bias.detach())
        return wc, bc, wd, bd

    def grad_scratch(self, dev):
        """(gw_c, gb_c, gw_d, gb_d) as views of ONE zeroed buffer: one fill per step instead of four."""
        sizes = (self.width_c * 128, self.width_c, self.width_d * 128, self.width_d)
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        a, b, c, d = torch.split(flat, sizes)
        return a.view(self.width_c, 128), b, c.view(self.width_d, 128), d


class _FusedHeadsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cen, den, n_vis, fh: FusedHeads, precision):
        from .dense import pack_weight, tc_linear
        cen, den = cen.contiguous(), den.contiguous()
        wc, bc, wd, bd = fh.cat_params(cen.device)
        pc, pd = pack_weight(wc, precision == 3), pack_weight(wd, precision == 3)
        out_c = tc_linear(cen[n_vis:], wc, n_out=fh.width_c, bias=bc, precision=precision, packed=pc)
        out_d = tc_linear(den[n_vis:], wd, n_out=fh.width_d, bias=bd, precision=precision, packed=pd)
        ctx.save_for_backward(cen, den)
        ctx.keep = (n_vis, fh, precision, wc, wd, pc, pd)
        return out_c, out_d

    @staticmethod
    def backward(ctx, d_c, d_d):
        from .dense import tc_linear, tc_wgrad
        cen, den = ctx.saved_tensors
        n_vis, fh, precision, wc, wd, pc, pd = ctx.keep
        d_c, d_d = d_c.contiguous(), d_d.contiguous()
        d_cen, d_den = torch.zeros_like(cen), torch.zeros_like(den)
        tc_linear(d_c, wc, n_out=128, w_mn_major=True, out=d_cen[n_vis:], precision=precision, packed=pc)
        tc_linear(d_d, wd, n_out=128, w_mn_major=True, out=d_den[n_vis:], precision=precision, packed=pd)
        gw_c, gb_c, gw_d, gb_d = fh.grad_scratch(cen.device)
        tc_wgrad(d_c, cen[n_vis:], gw_c, gb_c, precision=precision)
        tc_wgrad(d_d, den[n_vis:], gw_d, gb_d, precision=precision)
        dst, src = [], []
        for h, off in zip(fh.heads_c, fh.offsets):
            for p, g in ((h.weight, gw_c[off:off + h.out_features]), (h.bias, gb_c[off:off + h.out_features])):
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                dst.append(p.grad)
                src.append(g)
        k = fh.head_d.out_features
        for p, g in ((fh.head_d.weight, gw_d[:k]), (fh.head_d.bias, gb_d[:k])):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            dst.append(p.grad)
            src.append(g)
        torch._foreach_add_(dst, src)
        return d_cen, d_den, None, None, None


class _DecoderTokensFn(torch.autograd.Function):
    """[visible ; mask_token x n_mask] and its backward as one launch each (csrc/tokens.cu); the mask token's gradient
    is accumulated straight into ``mask_token.grad`` like every other parameter gradient of the fused path."""

    @staticmethod
    def forward(ctx, vis, mask_token, n_mask):
        from . import lib as L
        vis = vis.contiguous()
        n_vis, d = vis.shape
        out = torch.empty((n_vis + n_mask, d), dtype=torch.float32, device=vis.device)
        L.run("decoder_tokens", L.ptr(vis), n_vis, L.ptr(mask_token), n_mask, d, L.ptr(out), L.stream_ptr(vis.device))
        ctx.mask_token, ctx.shape = mask_token, (n_vis, n_mask, d)
        return out

    @staticmethod
    def backward(ctx, d_tokens):
        from . import lib as L
        from .sst import _grad_of
        n_vis, n_mask, d = ctx.shape
        d_tokens = d_tokens.contiguous()
        L.run("mask_token_grad", L.ptr(d_tokens), n_vis, n_mask, d, L.ptr(_grad_of(ctx.mask_token)),
              L.stream_ptr(d_tokens.device))
        return d_tokens[:n_vis], None, None


@BACKBONES.register_module()
class MultiMAESSTSPChoose(nn.Module):
    def __init__(self, window_shape, shifts_list, point_cloud_range, voxel_size, shuffle_voxels=False, d_model=[],
                 nhead=[], sub_voxel_ratio_low=[], sub_voxel_ratio_med=[], cls_sub_voxel=False, encoder_num_blocks=6,
                 decoder_num_blocks=2, dim_feedforward=[], dropout=0.0, activation="gelu", output_shape=None,
                 low=True, med=True, top=True, debug=True, drop_info=None, normalize_pos=False, pos_temperature=10000,
                 in_channel=None, conv_kwargs=None, checkpoint_blocks=[]):
        super().__init__()
        assert drop_info is not None
        if shuffle_voxels or normalize_pos or in_channel is not None:
            raise NotImplementedError("shuffle_voxels / normalize_pos / in_channel are off the GeoMAE path")
        assert len(set(d_model)) == 1, "one d_model for every block (…top_only.py:378)"
        self.window_shape, self.shifts_list = tuple(window_shape), list(shifts_list)
        self.point_cloud_range, self.voxel_size = point_cloud_range, voxel_size
        self.meta_drop_info, self.pos_temperature = drop_info, pos_temperature
        self.d_model, self.nhead = d_model, nhead
        self.cls_sub_voxel, self.low, self.med, self.top = cls_sub_voxel, low, med, top
        self.output_shape, self.debug = output_shape, debug
        max_tokens = max(v["max_tokens"] for v in (drop_info[0] if isinstance(drop_info, tuple) else drop_info).values())
        if window_shape[0] * window_shape[1] > max_tokens:
            raise NotImplementedError("token dropping (window larger than the largest bucket) is not on this path")
        self.spec = WindowSpec(window_shape, shifts_list)
        # sub-voxel sizes do not matter for window geometry; reuse the pillar size for all three scales
        self.geom = VoxelGeometry(tuple(point_cloud_range), tuple(voxel_size), tuple(voxel_size), tuple(voxel_size),
                                  (1, 1, 1), (1, 1, 1))

        def blocks(n):
            return nn.ModuleList([BasicShiftBlock(d_model[i], nhead[i], dim_feedforward[i], dropout, activation,
                                                  batch_first=False, block_id=i) for i in range(n)])
        self.encoder_blocks = blocks(encoder_num_blocks)
        self.decoder_centroid_blocks = blocks(decoder_num_blocks)
        self.decoder_density_blocks = blocks(decoder_num_blocks)
        d = d_model[-1]
        self.mask_token = nn.Parameter(torch.zeros(1, d))
        self.per_sub_voxel_num_low = sub_voxel_ratio_low[0] * sub_voxel_ratio_low[1] * sub_voxel_ratio_low[2]
        self.per_sub_voxel_num_med = sub_voxel_ratio_med[0] * sub_voxel_ratio_med[1] * sub_voxel_ratio_med[2]
        self.decoder_pred_low = nn.Linear(d, self.per_sub_voxel_num_low * 3)
        self.decoder_pred_med = nn.Linear(d, self.per_sub_voxel_num_med * 3)
        self.decoder_pred_top = nn.Linear(d, 3)
        if low:
            self.decoder_pred_density_low = nn.Linear(d, self.per_sub_voxel_num_low * 3)
        if med:
            self.decoder_pred_density_med = nn.Linear(d, self.per_sub_voxel_num_med * 3)
        if top:
            self.decoder_pred_density_top = nn.Linear(d, 3)
        if cls_sub_voxel:
            self.cls_pred_low = nn.Linear(d, self.per_sub_voxel_num_low * 2)
            self.cls_pred_med = nn.Linear(d, self.per_sub_voxel_num_med * 2)
        self._reset_parameters()

    def set_sra_impl(self, impl: str):
        """"tc3" (default, bf16x3 tensor-core, fp32 parity) | "tc1" (plain bf16 tensor-core) | "glue" (library GEMMs)."""
        assert impl in ("tc3", "tc1", "glue")
        for m in self.modules():
            if hasattr(m, "impl") and hasattr(m, "win_attn"):
                m.impl = impl
        self.sra_impl = impl

    def _precision(self):
        return 1 if getattr(self, "sra_impl", "tc3") == "tc1" else 3

    def _stack(self, blocks):
        """All EncoderLayers of a ModuleList of BasicShiftBlocks as one executor (layer j of a block uses shift j)."""
        key = id(blocks)
        cache = self.__dict__.setdefault("_stacks", {})
        if key not in cache:
            layers = [layer for block in blocks for layer in block.encoder_list]
            shifts = [j % self.spec.n_shifts for block in blocks for j in range(len(block.encoder_list))]
            cache[key] = SRAStack(layers, shifts)
        return cache[key]

    def _pos(self, layout):
        """(per-shift gathered position rows for the glue path | None, the [144,128] table for the fused path)."""
        table = pos_table(self.window_shape, self.d_model[0], self.pos_temperature, layout.tok_cell.device)
        if getattr(self, "sra_impl", "tc3") == "glue":
            return window_pos_embed(layout, self.d_model[0], self.pos_temperature), table
        return None, table

    def _reset_parameters(self):
        for name, p in self.named_parameters():     # …top_only.py:318-321
            if p.dim() > 1 and "scaler" not in name:
                nn.init.xavier_uniform_(p)

    def _layout(self, coors, batch_size, pillar_batch: PillarBatch | None, rows):
        if pillar_batch is not None:
            return WindowLayout.from_pillars(self.spec, pillar_batch, rows)
        return WindowLayout.from_coors(self.spec, self.geom, coors, batch_size)

    def build_layouts(self, pillar_batch: PillarBatch, rows_keep, rows_mask):
        """(encoder layout over the visible pillars, decoder layout over visible + masked) — index work that depends on
        the scatter result and the mask split only, so the detector runs it on the input stream under the VFE."""
        return (WindowLayout.from_pillars(self.spec, pillar_batch, rows_keep),
                WindowLayout.from_pillars(self.spec, pillar_batch, torch.cat([rows_keep, rows_mask])))

    def forward(self, voxel_feat, coors, coors_mask, batch_size, pillar_batch=None, rows_keep=None, rows_mask=None,
                layouts=None):
        """Reference call form ``backbone(voxel_feat, coors, coors_mask, batch_size)``; when the
        caller also holds the scatter result it passes it (with the pillar rows) to reuse its bitmap, and possibly
        the two window layouts it already built from it (``build_layouts``)."""
        enc_layout = layouts[0] if layouts else self._layout(coors, batch_size, pillar_batch, rows_keep)
        x = self.forward_encoder(voxel_feat, enc_layout)
        return self.forward_decoder(x, coors, coors_mask, batch_size, pillar_batch, rows_keep, rows_mask,
                                    layout=layouts[1] if layouts else None)

    def forward_encoder(self, voxel_feat, layout):
        pos, table = self._pos(layout)
        if pos is None:
            return self._stack(self.encoder_blocks)(voxel_feat, layout, table, self._precision())
        out = voxel_feat
        for block in self.encoder_blocks:
            out = block(out, layout, pos, table)
        return out

    def forward_decoder(self, visible_voxel_feat, coors, coors_mask, batch_size, pillar_batch=None, rows_keep=None,
                        rows_mask=None, layout=None):
        n_vis = visible_voxel_feat.shape[0]
        n_mask = coors_mask.shape[0] if coors_mask is not None else rows_mask.shape[0]
        hook = getattr(self, "encoder_output_hook", None)
        if hook is not None and visible_voxel_feat.requires_grad:
            # fires in backward once every decoder / head gradient has been queued (FlatTrainer: early all-reduce bucket)
            visible_voxel_feat.register_hook(lambda g: (hook(g), None)[1])
        if getattr(self, "sra_impl", "tc3") != "glue" and self.d_model[0] == 128 and visible_voxel_feat.is_cuda:
            tokens = _DecoderTokensFn.apply(visible_voxel_feat, self.mask_token, n_mask)
        else:
            tokens = torch.cat([visible_voxel_feat, self.mask_token.repeat(n_mask, 1)], dim=0)
        if layout is None:
            all_coors = torch.cat([coors, coors_mask], dim=0)
            rows = torch.cat([rows_keep, rows_mask]) if pillar_batch is not None else None
            layout = self._layout(all_coors, batch_size, pillar_batch, rows)
        pos, table = self._pos(layout)
        cen = den = tokens
        if pos is None:
            from .sst import _SRADualStackFn
            cen, den = _SRADualStackFn.apply(tokens, self._stack(self.decoder_centroid_blocks),
                                             self._stack(self.decoder_density_blocks), layout, table, self._precision())
        else:
            for block in self.decoder_centroid_blocks:
                cen = block(cen, layout, pos, table)
            for block in self.decoder_density_blocks:
                den = block(den, layout, pos, table)
        self._fused_heads = None
        if pos is None and self.cls_sub_voxel and self.top and not self.low and not self.med:
            # the configuration of configs/mae_sst/*: all six heads as two tensor-core GEMMs
            fh = self.__dict__.get("_fh") or self.__dict__.setdefault("_fh", FusedHeads(self))
            out_c, out_d = _FusedHeadsFn.apply(cen, den, n_vis, fh, self._precision())
            sl, sm = self.per_sub_voxel_num_low, self.per_sub_voxel_num_med
            o = fh.offsets
            self._fused_heads = (out_c, out_d, fh)
            return (out_c[:, o[0]:o[0] + sl * 3].view(-1, sl, 3), out_c[:, o[2]:o[2] + sm * 3].view(-1, sm, 3),
                    out_c[:, o[4]:o[4] + 3], None, None, out_d[:, :3],
                    out_c[:, o[1]:o[1] + sl * 2].view(-1, sl, 2), out_c[:, o[3]:o[3] + sm * 2].view(-1, sm, 2))
        cen, den = cen[n_vis:], den[n_vis:]
        reg_low = self.decoder_pred_low(cen).view(-1, self.per_sub_voxel_num_low, 3)
        reg_med = self.decoder_pred_med(cen).view(-1, self.per_sub_voxel_num_med, 3)
        reg_top = self.decoder_pred_top(cen)
        nor_low = self.decoder_pred_density_low(den).view(-1, self.per_sub_voxel_num_low, 3) if self.low else None
        nor_med = self.decoder_pred_density_med(den).view(-1, self.per_sub_voxel_num_med, 3) if self.med else None
        nor_top = self.decoder_pred_density_top(den) if self.top else None
        if self.cls_sub_voxel:
            cls_low = self.cls_pred_low(cen).view(-1, self.per_sub_voxel_num_low, 2)
            cls_med = self.cls_pred_med(cen).view(-1, self.per_sub_voxel_num_med, 2)
            return reg_low, reg_med, reg_top, nor_low, nor_med, nor_top, cls_low, cls_med
        return reg_low, reg_med, reg_top, nor_low, nor_med, nor_top
