"""SSTSecondPretrainedv1 — mirror of mmdet3d/models/backbones/sst_second_pretrained_v1.py:17-233 (same registry key,
constructor arguments and module tree, so ``encoder_blocks.*`` of a GeoMAE pre-training checkpoint load by key and
``conv_blocks.*`` keep the reference's names): the SST encoder of the pre-training path followed by the dense BEV
canvas and three SECOND-style convolution stages.  SURVEY.md §8(f) N1.

The encoder runs on the same fused SRA executor as the pre-training backbone (csrc/sra_stack.cu, CSR windows from
``voxel_info['window_layout']``); ``recover_bev`` is one kernel (csrc/bev.cu).  The convolution stages are plain
library convolutions (cuDNN through ``nn.Conv2d``) — they are outside the pre-training hot path and are kept only so
the consumer is complete."""
from __future__ import annotations

import torch
from torch import nn

from . import lib as L
from .registry import BACKBONES, build_norm_layer
from .sst import BasicShiftBlock, SRAStack
from .voxel import VoxelGeometry
from .windows import WindowLayout, WindowSpec, pos_table


class _RecoverBEV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, coors32, batch_size, ny, nx):
        feat = feat.contiguous()
        n, c = feat.shape
        canvas = torch.empty((batch_size, c, ny, nx), dtype=torch.float32, device=feat.device)
        L.run("recover_bev", L.ptr(feat), L.ptr(coors32), n, c, batch_size, ny, nx, L.ptr(canvas),
              L.stream_ptr(feat.device))
        ctx.save_for_backward(coors32)
        ctx.shape = (n, c, ny, nx)
        return canvas

    @staticmethod
    def backward(ctx, d_canvas):
        (coors32,) = ctx.saved_tensors
        n, c, ny, nx = ctx.shape
        d_canvas = d_canvas.contiguous()
        d_feat = torch.empty((n, c), dtype=torch.float32, device=d_canvas.device)
        L.run("recover_bev_bwd", L.ptr(d_canvas), L.ptr(coors32), n, c, ny, nx, L.ptr(d_feat),
              L.stream_ptr(d_canvas.device))
        return d_feat, None, None, None, None


@BACKBONES.register_module()
class SSTSecondPretrainedv1(nn.Module):
    def __init__(self, eval_flag=False, model_path="", d_model=[], nhead=[], num_blocks=6, dim_feedforward=[],
                 dropout=0.0, activation="gelu", output_shape=None, num_attached_conv=2, conv_in_channels=64,
                 conv_out_channels=[128, 128, 256], layer_nums=[3, 5, 5], layer_strides=[2, 2, 2],
                 norm_cfg=dict(type="naiveSyncBN2d", eps=1e-3, momentum=0.01), conv_cfg=dict(type="Conv2d", bias=False),
                 debug=True, drop_info=None, normalize_pos=False, pos_temperature=10000, window_shape=None,
                 in_channel=None, conv_kwargs=dict(kernel_size=3, dilation=2, padding=2, stride=1),
                 checkpoint_blocks=[], shifts_list=None):
        super().__init__()
        assert drop_info is not None
        if normalize_pos:
            raise NotImplementedError("normalize_pos is off in every GeoMAE config")
        if conv_cfg.get("type", "Conv2d") != "Conv2d":
            raise NotImplementedError(f"conv_cfg type {conv_cfg['type']}")
        assert len(set(d_model)) == 1, "one d_model for every block (sst_second_pretrained_v1.py:296-297)"
        self.meta_drop_info, self.pos_temperature = drop_info, pos_temperature
        self.d_model, self.nhead, self.window_shape = d_model, nhead, tuple(window_shape)
        self.normalize_pos, self.checkpoint_blocks = normalize_pos, checkpoint_blocks   # activations are saved in bf16, never recomputed
        self.output_shape, self.debug = output_shape, debug
        # the reference reads the shifts from the input tuple's length; the CSR layout carries them
        self._shifts_list = shifts_list
        if in_channel is not None:
            self.linear0 = nn.Linear(in_channel, d_model[0])
        self.encoder_blocks = nn.ModuleList([
            BasicShiftBlock(d_model[i], nhead[i], dim_feedforward[i], dropout, activation, batch_first=False, block_id=i)
            for i in range(num_blocks)])
        self._reset_parameters()
        bias = conv_cfg.get("bias", False)
        in_filters = [conv_in_channels, *conv_out_channels[:-1]]
        stages = []
        for i, layer_num in enumerate(layer_nums):      # :137-166
            stage = [nn.Conv2d(in_filters[i], conv_out_channels[i], 3, stride=layer_strides[i], padding=1, bias=bias),
                     build_norm_layer(norm_cfg, conv_out_channels[i])[1], nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                stage += [nn.Conv2d(conv_out_channels[i], conv_out_channels[i], 3, padding=1, bias=bias),
                          build_norm_layer(norm_cfg, conv_out_channels[i])[1], nn.ReLU(inplace=True)]
            stages.append(nn.Sequential(*stage))
        self.conv_blocks = nn.ModuleList(stages)
        self.sra_impl = "tc3"

    def set_sra_impl(self, impl: str):
        assert impl in ("tc3", "tc1")
        self.sra_impl = impl

    def _reset_parameters(self):
        for name, p in self.named_parameters():     # :238-241
            if p.dim() > 1 and "scaler" not in name:
                nn.init.xavier_uniform_(p)

    def set_drop_info(self):
        if hasattr(self, "drop_info"):
            return
        meta = self.meta_drop_info
        self.drop_info = (meta[0] if self.training else meta[1]) if isinstance(meta, tuple) else meta

    def _stack(self, n_shifts):
        st = self.__dict__.get("_sra_stack")
        if st is None:
            layers = [layer for block in self.encoder_blocks for layer in block.encoder_list]
            shifts = [j % n_shifts for block in self.encoder_blocks for j in range(len(block.encoder_list))]
            st = self.__dict__["_sra_stack"] = SRAStack(layers, shifts)
        return st

    def _layout_of(self, ind_dict_list, voxel_info):
        layout = voxel_info.get("window_layout")
        if layout is not None:
            return layout
        # input produced by other code (reference-format tuple): rebuild the CSR windows from the coordinates
        n_shifts = len(ind_dict_list)
        wx, wy = self.window_shape
        shifts = self._shifts_list or [(0, 0), (wx // 2, wy // 2)][:n_shifts]
        coors = voxel_info["coors"]
        batch_size = int(coors[:, 0].max().item()) + 1
        ny, nx = self.output_shape
        # only the grid extent matters for window geometry: unit pillars over [0,nx) x [0,ny)
        geom = VoxelGeometry((0.0, 0.0, 0.0, float(nx), float(ny), 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0),
                             (1.0, 1.0, 1.0), (1, 1, 1), (1, 1, 1))
        return WindowLayout.from_coors(WindowSpec(self.window_shape, shifts), geom, coors, batch_size)

    def forward(self, input_tuple):
        """(voxel_feat, ind_dict_list, voxel_info) of SSTInputLayer -> tuple of the three stage outputs (:170-214)."""
        voxel_feat, ind_dict_list, voxel_info = input_tuple
        coors = voxel_info["coors"]
        assert coors.dtype == torch.int64, "data type of coors should be torch.int64!"
        self.set_drop_info()
        layout = self._layout_of(ind_dict_list, voxel_info)
        batch_size = layout.n_frames
        output = voxel_feat
        if hasattr(self, "linear0"):
            output = self.linear0(output)
        table = pos_table(self.window_shape, self.d_model[0], self.pos_temperature, output.device)
        output = self._stack(layout.spec.n_shifts)(output, layout, table, 1 if self.sra_impl == "tc1" else 3)
        ny, nx = self.output_shape
        output = _RecoverBEV.apply(output, coors.to(torch.int32).contiguous(), batch_size, ny, nx)
        outs = []
        # parity mode keeps the library convolutions in fp32 as well (cuDNN would otherwise pick TF32); their backward
        # follows torch.backends.cudnn.allow_tf32, which the caller owns
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=self.sra_impl == "tc1"):
            for stage in self.conv_blocks:
                output = stage(output)
                outs.append(output)
        return tuple(outs)

    def recover_bev(self, voxel_feat, coors, batch_size):
        """:246-276, one launch for the batch."""
        ny, nx = self.output_shape
        return _RecoverBEV.apply(voxel_feat, coors.to(torch.int32).contiguous(), batch_size, ny, nx)
