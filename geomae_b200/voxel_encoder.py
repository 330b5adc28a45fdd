"""DynamicScatterVFE — mirror of mmdet3d/models/voxel_encoders/voxel_encoder.py:308-419 and
DynamicVFELayer (voxel_encoders/utils.py:107-144), same registry key and state_dict keys.

The reference re-derives the point->pillar map with torch.unique(dim=0) for every scatter
(3 sorts per forward); here it is an input (``PillarBatch.point_pillar``) produced once by the
fused scatter stage, and scatter-mean/max run as single-pass kernels (csrc/vfe.cu)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import lib as L
from .registry import VOXEL_ENCODERS, build_norm_layer
from .voxel import PillarBatch


class _ScatterReduce(torch.autograd.Function):
    MODES = {"sum": 0, "mean": 1, "avg": 1, "max": 2}

    @staticmethod
    def forward(ctx, feat, point_pillar, pillar_mean, n_pillars, mode):
        feat = feat.contiguous()
        n, c = feat.shape
        out = torch.empty((n_pillars, c), dtype=torch.float32, device=feat.device)
        arg = torch.empty((n_pillars, c), dtype=torch.int32, device=feat.device) if mode == 2 else None
        L.run("scatter_reduce_fwd", L.ptr(feat), n, c, L.ptr(point_pillar), L.ptr(pillar_mean),
                                                  n_pillars, mode, L.ptr(out), L.ptr(arg),
                                                  L.stream_ptr(feat.device))
        ctx.mode, ctx.shape = mode, (n, c)
        ctx.save_for_backward(point_pillar, pillar_mean, arg)
        return out

    @staticmethod
    def backward(ctx, d_out):
        point_pillar, pillar_mean, arg = ctx.saved_tensors
        n, c = ctx.shape
        d_out = d_out.contiguous()
        d_feat = torch.empty((n, c), dtype=torch.float32, device=d_out.device)
        L.run("scatter_reduce_bwd", L.ptr(d_out), n, c, L.ptr(point_pillar), L.ptr(pillar_mean),
                                                  L.ptr(arg), ctx.mode, L.ptr(d_feat), L.stream_ptr(d_out.device))
        return d_feat, None, None, None, None


def scatter_reduce(feat, pb: PillarBatch, mode: str):
    """Reduce per-point rows into per-pillar rows (lexicographic pillar order)."""
    L.require_cuda(feat, "feat")
    return _ScatterReduce.apply(feat, pb.point_pillar, pb.pillar_mean, pb.n_pillars, _ScatterReduce.MODES[mode])


class DynamicVFELayer(nn.Module):
    """Linear(no bias) -> norm -> ReLU (utils.py:118-144)."""

    def __init__(self, in_channels, out_channels, norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01)):
        super().__init__()
        self.norm = build_norm_layer(norm_cfg, out_channels)[1]
        self.linear = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, inputs, tc_precision=0):
        lin = self.linear
        if tc_precision and lin.in_features % 128 == 0 and lin.out_features % 128 == 0:
            from .dense import tc_linear_module
            x = tc_linear_module(inputs, lin, tc_precision)      # tcgen05 GEMM (the 128->128 layer: 3.6 GFLOP/step)
        else:
            x = lin(inputs)
        return F.relu(self.norm(x))


@VOXEL_ENCODERS.register_module()
class DynamicScatterVFE(nn.Module):
    def __init__(self, in_channels=4, feat_channels=[], with_distance=False, with_cluster_center=False,
                 with_voxel_center=False, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), mode="max", fusion_layer=None,
                 return_point_feats=False, return_inv=True, rel_dist_scaler=1.0, unique_once=False):
        super().__init__()
        assert mode in ("avg", "max") and len(feat_channels) > 0
        if with_distance or fusion_layer is not None or return_point_feats:
            raise NotImplementedError("with_distance / fusion_layer / return_point_feats are off the GeoMAE path")
        if not (with_cluster_center and with_voxel_center):
            raise NotImplementedError("the fused decoration kernel emits cluster and voxel-centre offsets")
        if rel_dist_scaler != 1.0:
            raise NotImplementedError("rel_dist_scaler != 1")
        self.raw_channels = in_channels
        self.in_channels = in_channels + 6
        self.mode = mode
        self.vx, self.vy, self.vz = voxel_size
        # voxel_encoder.py:155-157 — evaluated in Python floats, rounded to fp32 when they meet the tensor
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.z_offset = self.vz / 2 + point_cloud_range[2]
        self.point_cloud_range = point_cloud_range
        chans = [self.in_channels] + list(feat_channels)
        self.vfe_layers = nn.ModuleList([
            DynamicVFELayer(chans[i] * (2 if i > 0 else 1), chans[i + 1], norm_cfg) for i in range(len(chans) - 1)])
        self.num_vfe = len(self.vfe_layers)
        self.tc_precision = 3       # 0: library GEMM, 1: bf16 tensor-core, 3: bf16x3 tensor-core (fp32 parity)

    def decorate(self, pb: PillarBatch):
        pts = pb.points
        n, c = pts.shape
        if c != self.raw_channels:
            raise RuntimeError(f"points have {c} channels, encoder was built for {self.raw_channels}")
        out = torch.empty((n, c + 6), dtype=torch.float32, device=pts.device)
        L.run("vfe_decorate", L.ptr(pts), n, c, L.ptr(pb.point_pillar), L.ptr(pb.pillar_mean),
                                            L.ptr(pb.pillar_coors), L.f3((self.vx, self.vy, self.vz)),
                                            L.f3((self.x_offset, self.y_offset, self.z_offset)), L.ptr(out),
                                            L.stream_ptr(pts.device))
        return out

    def forward(self, pb: PillarBatch, return_inv=False):
        """-> voxel_feats [V, C_out], voxel_coors [V,4] int32 (b,z,y,x) sorted (, point->pillar map)."""
        features = self.decorate(pb)
        voxel_feats = None
        for i, vfe in enumerate(self.vfe_layers):
            point_feats = vfe(features, self.tc_precision)
            voxel_feats = scatter_reduce(point_feats, pb, self.mode)
            if i != self.num_vfe - 1:
                features = torch.cat([point_feats, voxel_feats.index_select(0, pb.point_pillar.long())], dim=1)
        coors = pb.pillar_coors[:pb.n_pillars]
        if return_inv:
            return voxel_feats, coors, pb.point_pillar
        return voxel_feats, coors
