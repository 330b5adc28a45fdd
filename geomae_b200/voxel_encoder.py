"""DynamicScatterVFE — mirror of mmdet3d/models/voxel_encoders/voxel_encoder.py:308-419 and
DynamicVFELayer (voxel_encoders/utils.py:107-144), same registry key and state_dict keys.

The reference re-derives the point->pillar map with torch.unique(dim=0) for every scatter
(3 sorts per forward); here it is an input (``PillarBatch.point_pillar``) produced once by the
fused scatter stage, and scatter-mean/max run as single-pass kernels (csrc/vfe.cu)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import distributed as dist
from torch import nn

from . import lib as L
from .norm import NaiveSyncBatchNorm1d
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


def _bn_moments(stats, n, world):
    """[sum | sum of squares] (fp64) of this rank's n rows -> [E x | E x^2] averaged over ranks with EQUAL weight per
    rank (the rule of naiveSyncBN1d, mmdet3d/ops/norm.py:66-73), in place.  One launch: the normalisation by n and —
    with several ranks on one node — the exchange itself over peer memory (csrc/peer.cu); ``torch.distributed`` only
    when the peer mailboxes could not be set up."""
    from . import peer
    if world == 1:
        return peer.scale_(stats, 1.0 / float(max(n, 1)))
    px = peer.PeerExchange.get(stats.device)
    if px is not None:
        return px.allreduce_(stats, pre_scale=1.0 / float(max(n, 1)), post_scale=1.0 / world)
    mom = stats / float(max(n, 1))
    dist.all_reduce(mom)
    mom /= world
    return mom


def _sum_over_ranks(buf, world):
    """In-place sum of a small fp64 vector over the ranks (BatchNorm backward coefficients)."""
    if world > 1:
        from . import peer
        px = peer.PeerExchange.get(buf.device)
        if px is not None:
            px.allreduce_(buf)
        else:
            dist.all_reduce(buf)
    return buf


class _FusedVFEFn(torch.autograd.Function):
    """The whole two-layer DynamicScatterVFE (max mode, BatchNorm in training mode) on the fused kernels of
    csrc/vfe_fused.cu + one tcgen05 GEMM.  Parameter gradients are accumulated straight into ``p.grad``; the only
    autograd input is a parameter used as an anchor (raw points carry no gradient)."""

    @staticmethod
    def forward(ctx, anchor, vfe, pb, precision):
        from .dense import tc_linear
        l0, l1 = vfe.vfe_layers
        dev = pb.points.device
        n, v = pb.points.shape[0], pb.n_pillars
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        sync = world > 1 and isinstance(l0.norm, NaiveSyncBatchNorm1d)
        world = world if sync else 1
        s = L.stream_ptr(dev)
        f32, f64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.float64, device=dev)
        vox, off = L.f3((vfe.vx, vfe.vy, vfe.vz)), L.f3((vfe.x_offset, vfe.y_offset, vfe.z_offset))
        unbias = 1.0 if sync else n / max(n - 1, 1)

        def bn_args(layer):
            nm = layer.norm
            return L.ptr(nm.weight), L.ptr(nm.bias), float(nm.eps)

        w0, w1 = l0.linear.weight, l1.linear.weight
        x1, stats0 = torch.empty((n, 64), **f32), torch.empty(128, **f64)
        L.run("vfe0_forward", L.ptr(pb.points), n, pb.points.shape[1], L.ptr(pb.point_pillar), L.ptr(pb.pillar_mean),
              L.ptr(pb.pillar_coors), vox, off, L.ptr(w0), L.ptr(x1), L.ptr(stats0), s)
        mom0 = _bn_moments(stats0, n, world)
        vmax1 = torch.empty((max(v, 1), 64), dtype=torch.int64, device=dev)
        L.run("vfe_bn_relu_max", L.ptr(x1), n, 64, L.ptr(pb.point_pillar), L.ptr(mom0), *bn_args(l0),
              L.ptr(l0.norm.running_mean), L.ptr(l0.norm.running_var), float(l0.norm.momentum), unbias, L.ptr(vmax1), v, s)
        feat1 = torch.empty((n, 128), **f32)
        L.run("vfe_cat", L.ptr(x1), n, L.ptr(pb.point_pillar), L.ptr(mom0), *bn_args(l0), L.ptr(vmax1), L.ptr(feat1), s)
        x2 = tc_linear(feat1, w1, n_out=128, precision=precision)
        stats1 = torch.empty(256, **f64)
        L.run("colstats", L.ptr(x2), n, 128, L.ptr(stats1), s)
        mom1 = _bn_moments(stats1, n, world)
        vmax2 = torch.empty((max(v, 1), 128), dtype=torch.int64, device=dev)
        L.run("vfe_bn_relu_max", L.ptr(x2), n, 128, L.ptr(pb.point_pillar), L.ptr(mom1), *bn_args(l1),
              L.ptr(l1.norm.running_mean), L.ptr(l1.norm.running_var), float(l1.norm.momentum), unbias, L.ptr(vmax2), v, s)
        out = torch.empty((v, 128), **f32)
        L.run("vmax_decode", L.ptr(vmax2), v * 128, L.ptr(out), s)
        if not sync:
            torch._foreach_add_([l0.norm.num_batches_tracked, l1.norm.num_batches_tracked], 1)
        ctx.keep = (vfe, pb, precision, world, x1, feat1, x2, mom0, mom1, vmax1, vmax2, vox, off)
        return out

    @staticmethod
    def backward(ctx, d_vox):
        from .dense import tc_linear, tc_wgrad
        vfe, pb, precision, world, x1, feat1, x2, mom0, mom1, vmax1, vmax2, vox, off = ctx.keep
        l0, l1 = vfe.vfe_layers
        dev = d_vox.device
        n, v = pb.points.shape[0], pb.n_pillars
        s = L.stream_ptr(dev)
        f32, f64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.float64, device=dev)
        d_vox = d_vox.contiguous()
        inv_wn = 1.0 / (world * max(n, 1))

        def grad(p):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            return p.grad

        def bn_args(layer):
            nm = layer.norm
            return L.ptr(nm.weight), L.ptr(nm.bias), float(nm.eps)

        def coeffs(layer, c, sums, mom):
            ab = torch.empty(2 * c, **f64)
            nm = layer.norm
            L.run("bn_backward_coeffs", c, L.ptr(sums), L.ptr(mom), L.ptr(nm.weight), float(nm.eps), L.ptr(ab),
                  L.ptr(grad(nm.weight)), L.ptr(grad(nm.bias)), s)
            return _sum_over_ranks(ab, world)

        # ---- layer 1
        sums1 = torch.empty(256, **f64)
        common1 = (L.ptr(x2), n, L.ptr(pb.point_pillar), L.ptr(mom1), *bn_args(l1), L.ptr(vmax2), L.ptr(d_vox))
        L.run("vfe1_backward", 0, *common1, L.ptr(sums1), None, 0.0, None, s)
        ab1 = coeffs(l1, 128, sums1, mom1)
        dx2 = torch.empty((n, 128), **f32)
        L.run("vfe1_backward", 1, *common1, None, L.ptr(ab1), inv_wn, L.ptr(dx2), s)
        w1 = l1.linear.weight
        tc_wgrad(dx2, feat1, grad(w1), None, precision=precision)
        dfeat1 = tc_linear(dx2, w1, n_out=128, w_mn_major=True, precision=precision)
        # ---- pillar-max gather of layer 1's operand, then layer 0
        d_vmax1 = torch.empty((max(v, 1), 64), **f32)
        L.run("vfe_gather_backward", L.ptr(dfeat1), n, L.ptr(pb.point_pillar), L.ptr(d_vmax1), v, s)
        sums0 = torch.empty(128, **f64)
        common0 = (L.ptr(pb.points), n, pb.points.shape[1], L.ptr(pb.point_pillar), L.ptr(pb.pillar_mean),
                   L.ptr(pb.pillar_coors), vox, off, L.ptr(x1), L.ptr(mom0), *bn_args(l0), L.ptr(vmax1), L.ptr(d_vmax1),
                   L.ptr(dfeat1))
        L.run("vfe0_backward", 0, *common0, L.ptr(sums0), None, 0.0, None, s)
        ab0 = coeffs(l0, 64, sums0, mom0)
        L.run("vfe0_backward", 1, *common0, None, L.ptr(ab0), inv_wn, L.ptr(grad(l0.linear.weight)), s)
        return None, None, None, None


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

    def _can_fuse(self, pb):
        l = self.vfe_layers
        return (self.tc_precision in (1, 3) and self.training and self.mode == "max" and self.num_vfe == 2
                and self.raw_channels == 5 and pb.points.shape[1] == 5 and pb.points.shape[0] > 1
                and l[0].linear.out_features == 64 and l[1].linear.in_features == 128 and l[1].linear.out_features == 128
                and all(isinstance(x.norm, nn.BatchNorm1d) and x.norm.affine and x.norm.track_running_stats
                        and x.norm.momentum is not None for x in l)
                and getattr(self, "fused", True))

    def _pillars_from(self, features, coors):
        """PillarBatch for the reference call form ``forward(features [P,C], coors [P,4] (b,z,y,x))``
        (voxel_encoders/voxel_encoder.py:358-364): ``coors`` must be the dynamic voxelisation of ``features`` at this
        encoder's own voxel size / range (what the detector passes); only its batch column is read here — the fused
        scatter recomputes the coordinates bit-exactly and returns the same sorted pillar list as torch.unique(coors)."""
        from .voxel import VoxelGeometry
        L.require_cuda(features, "features")
        if coors.shape[0] != features.shape[0] or coors.shape[1] != 4:
            raise RuntimeError("DynamicScatterVFE: coors must be [P, 4] (batch, z, y, x) rows of the points")
        n_frames = int(coors[-1, 0]) + 1 if coors.shape[0] else 1          # compat path: one host read
        geom = self.__dict__.get("_geom")
        if geom is None:
            vs = (self.vx, self.vy, self.vz)
            geom = self.__dict__["_geom"] = VoxelGeometry(tuple(self.point_cloud_range), vs, vs, vs, (1, 1, 1), (1, 1, 1))
        counts = torch.bincount(coors[:, 0].long(), minlength=n_frames)
        offsets = torch.zeros(n_frames + 1, dtype=torch.int32, device=features.device)
        offsets[1:] = torch.cumsum(counts, 0).int()
        return PillarBatch(geom, features.float().contiguous(), offsets, n_frames).run()

    def forward(self, pb, coors=None, points=None, img_feats=None, img_metas=None, return_inv=False):
        """-> voxel_feats [V, C_out], voxel_coors [V,4] int32 (b,z,y,x) sorted (, point->pillar map).
        ``pb``: the PillarBatch of the fused scatter stage (hot path), or the reference's ``features`` tensor together
        with ``coors`` (voxel_encoder.py:358-364)."""
        if isinstance(pb, torch.Tensor):
            pb = self._pillars_from(pb, coors)
        if self._can_fuse(pb):       # the GeoMAE configuration in training: fused kernels (csrc/vfe_fused.cu)
            voxel_feats = _FusedVFEFn.apply(self.vfe_layers[0].linear.weight, self, pb, self.tc_precision)
            coors = pb.pillar_coors[:pb.n_pillars]
            return (voxel_feats, coors, pb.point_pillar) if return_inv else (voxel_feats, coors)
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
