"""MultiSubVoxelDynamicVoxelNetSSL — mirror of
mmdet3d/models/detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py (same registry key, ctor kwargs,
loss keys), restricted to the branch the shipped configs take: normalize_sub_voxel=True,
mse_loss=True, cls_sub_voxel=True, vanilla random masking."""
from __future__ import annotations

import torch
from torch import nn

from . import lib as L
from .registry import DETECTORS, build_backbone, build_loss, build_voxel_encoder
from .voxel import PillarBatch, VoxelGeometry, Voxelization, scatter_frames


class _GeomLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reg_low, reg_med, reg_top, nor_top, cls_low, cls_med, pb, rows, normal, weights):
        import ctypes as C
        preds = [t.contiguous() for t in (reg_low, reg_med, reg_top, nor_top, cls_low, cls_med)]
        rows = rows.to(torch.int64).contiguous()
        dev = preds[0].device
        a = L.LossArgs()
        a.rows, a.m = rows.data_ptr(), rows.shape[0]
        (a.reg_low, a.reg_med, a.reg_top, a.nor_top, a.cls_low, a.cls_med) = [p.data_ptr() for p in preds]
        a.normal = normal.data_ptr()
        a.w_low, a.w_med, a.w_top, a.w_nor, a.w_cls_low, a.w_cls_med = weights
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        acc = torch.empty(6, dtype=torch.float64, device=dev)
        out = torch.empty(6, dtype=torch.float32, device=dev)
        L.run("geom_loss_fwd", C.byref(pb.geom.cstruct), C.byref(pb.io), C.byref(a), L.ptr(counts), L.ptr(acc), L.ptr(out),
              L.stream_ptr(dev))
        ctx.args, ctx.pb, ctx.keep = a, pb, (preds, rows, normal, counts)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        preds, rows, normal, counts = ctx.keep
        g = g.contiguous()
        grads = [torch.empty_like(p) for p in preds]
        L.run("geom_loss_bwd", C.byref(ctx.pb.geom.cstruct), C.byref(ctx.pb.io), C.byref(ctx.args), L.ptr(counts), L.ptr(g),
              *[L.ptr(t) for t in grads], L.stream_ptr(g.device))
        return (*grads, None, None, None, None)


class _GeomLossFusedFn(torch.autograd.Function):
    """The same fused loss reading the predictions as column slices of the fused head outputs (row stride 768 / 128)."""

    @staticmethod
    def forward(ctx, out_c, out_d, fh, pb, rows, normal, weights):
        import ctypes as C
        rows = rows.to(torch.int64).contiguous()
        dev = out_c.device
        o = fh.offsets
        a = L.LossArgs()
        a.rows, a.m = rows.data_ptr(), rows.shape[0]
        base = out_c.data_ptr()
        a.reg_low, a.cls_low, a.reg_med, a.cls_med, a.reg_top = [base + 4 * k for k in o]
        a.nor_top = out_d.data_ptr()
        a.normal = normal.data_ptr()
        a.w_low, a.w_med, a.w_top, a.w_nor, a.w_cls_low, a.w_cls_med = weights
        wc, wd = out_c.shape[1], out_d.shape[1]
        for k, v in enumerate((wc, wc, wc, wd, wc, wc)):       # reg_low, reg_med, reg_top, nor_top, cls_low, cls_med
            a.ld[k] = v
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        acc = torch.empty(6, dtype=torch.float64, device=dev)
        out = torch.empty(6, dtype=torch.float32, device=dev)
        L.run("geom_loss_fwd", C.byref(pb.geom.cstruct), C.byref(pb.io), C.byref(a), L.ptr(counts), L.ptr(acc), L.ptr(out),
              L.stream_ptr(dev))
        ctx.args, ctx.pb, ctx.keep = a, pb, (out_c, out_d, rows, normal, counts, o)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        out_c, out_d, rows, normal, counts, o = ctx.keep
        g = g.contiguous()
        d_c, d_d = torch.zeros_like(out_c), torch.zeros_like(out_d)     # pad columns must carry zero gradient
        b = d_c.data_ptr()
        d_low, d_cls_low, d_med, d_cls_med, d_top = [b + 4 * k for k in o]
        L.run("geom_loss_bwd", C.byref(ctx.pb.geom.cstruct), C.byref(ctx.pb.io), C.byref(ctx.args), L.ptr(counts), L.ptr(g),
              d_low, d_med, d_top, d_d.data_ptr(), d_cls_low, d_cls_med, L.stream_ptr(g.device))
        return d_c, d_d, None, None, None, None, None


@DETECTORS.register_module()
class MultiSubVoxelDynamicVoxelNetSSL(nn.Module):
    def __init__(self, loss, loss_ratio_low, loss_ratio_med, loss_ratio_top, loss_ratio_low_nor, loss_ratio_med_nor,
                 loss_ratio_top_nor, hard_sub_voxel_layer_low, hard_sub_voxel_layer_med, hard_sub_voxel_layer_top,
                 random_mask_ratio, grid_size, sub_voxel_ratio_low, sub_voxel_ratio_med, voxel_layer,
                 sub_voxel_layer_low, sub_voxel_layer_med, voxel_encoder, backbone, spatial_shape=[1, 468, 468],
                 nor_usr_sml1=None, cls_loss_ratio_low=None, cls_loss_ratio_med=None, vis=False, cls_sub_voxel=False,
                 normalize_sub_voxel=None, use_focal_mask=None, norm_curv=True, mse_loss=None, neck=None,
                 bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        if neck is not None or bbox_head is not None:
            raise NotImplementedError("neck / bbox_head are not part of the SSL pretraining path")
        if use_focal_mask is not None or nor_usr_sml1 is not None or not mse_loss or not normalize_sub_voxel \
                or not cls_sub_voxel or not norm_curv:
            raise NotImplementedError("only the branch taken by configs/mae_sst/* is implemented "
                                      "(mse_loss, normalize_sub_voxel, cls_sub_voxel, vanilla mask)")
        self.backbone = build_backbone(backbone)
        self.voxel_encoder = build_voxel_encoder(voxel_encoder)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.spatial_shape, self.grid_size = spatial_shape, grid_size
        self.loss_ratio_low, self.loss_ratio_med, self.loss_ratio_top = loss_ratio_low, loss_ratio_med, loss_ratio_top
        self.loss_ratio_low_nor, self.loss_ratio_med_nor, self.loss_ratio_top_nor = \
            loss_ratio_low_nor, loss_ratio_med_nor, loss_ratio_top_nor
        self.cls_loss_ratio_low, self.cls_loss_ratio_med = cls_loss_ratio_low, cls_loss_ratio_med
        self.random_mask_ratio = random_mask_ratio
        self.point_cloud_range, self.voxel_size = voxel_layer["point_cloud_range"], voxel_layer["voxel_size"]
        self.sub_voxel_size_low, self.sub_voxel_size_med = sub_voxel_layer_low["voxel_size"], sub_voxel_layer_med["voxel_size"]
        self.sub_voxel_ratio_low, self.sub_voxel_ratio_med = sub_voxel_ratio_low, sub_voxel_ratio_med
        self.voxel_layer = Voxelization(**voxel_layer)
        self.sub_voxel_layer_low = Voxelization(**sub_voxel_layer_low)
        self.sub_voxel_layer_med = Voxelization(**sub_voxel_layer_med)
        # hard_sub_voxel_layer_* are constructed but never called by the reference (…_ssl.py:100-102); accepted, unused
        self.reg_loss = build_loss(loss)
        self.cls_loss = build_loss(dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0))
        self.geom = VoxelGeometry(tuple(self.point_cloud_range), tuple(self.voxel_size), tuple(self.sub_voxel_size_med),
                                  tuple(self.sub_voxel_size_low), tuple(sub_voxel_ratio_med), tuple(sub_voxel_ratio_low))
        gx, gy, gz = self.geom_grid_host()
        if (gz, gy, gx) != tuple(grid_size):
            raise ValueError(f"grid_size {tuple(grid_size)} does not match the voxel geometry {(gz, gy, gx)}")

    def set_impl(self, impl: str):
        """"tc3": tensor-core GEMMs in the bf16x3 split (fp32-equivalent, the parity mode; default);
        "tc1": plain bf16 tensor-core GEMMs (the bf16 benchmark mode); "glue": library fp32 GEMMs."""
        self.backbone.set_sra_impl(impl)
        self.voxel_encoder.tc_precision = {"tc3": 3, "tc1": 1, "glue": 0}[impl]
        return self

    def geom_grid_host(self):
        import math
        r, v = self.point_cloud_range, self.voxel_size
        return [int(math.ceil(round((r[3 + i] - r[i]) / v[i], 4))) for i in range(3)]

    # ------------------------------------------------------------------ reference-named pieces
    def voxelize(self, points):
        """…_ssl.py:355-377: per-sample coords with the batch index prepended."""
        coors = [nn.functional.pad(self.voxel_layer(p), (1, 0), value=i) for i, p in enumerate(points)]
        return torch.cat(points, dim=0), torch.cat(coors, dim=0)

    def sub_voxelize_low(self, points):
        return torch.cat([nn.functional.pad(self.sub_voxel_layer_low(p), (1, 0), value=i) for i, p in enumerate(points)])

    def sub_voxelize_med(self, points):
        return torch.cat([nn.functional.pad(self.sub_voxel_layer_med(p), (1, 0), value=i) for i, p in enumerate(points)])

    @torch.no_grad()
    def get_vanilla_mask_index(self, coors, batch_size, counts=None, frame_starts=None):
        """…_ssl.py:287-304: per sample keep int(L*(1-ratio)) random pillars, mask the rest.

        The reference draws torch.randperm(L) per sample; here ONE kernel (csrc/mask_split.cu) selects the same
        number of pillars uniformly at random per frame (seeded from torch's CPU generator) and returns both index
        lists in ascending pillar order — the model only depends on the subset.  ``counts`` (per-sample pillar
        counts on the host) and ``frame_starts`` (device int32 [B+1]) come for free from the scatter stage."""
        dev = coors.device
        if counts is None:
            counts = torch.bincount(coors[:, 0].long(), minlength=batch_size).tolist()
        if frame_starts is None:
            starts = [0]
            for n in counts:
                starts.append(starts[-1] + n)
            frame_starts = torch.tensor(starts, dtype=torch.int32, device=dev)
        keep_frac = 1 - self.random_mask_ratio
        n_keep = sum(int(n * keep_frac) for n in counts)
        ids_keep = torch.empty(n_keep, dtype=torch.int64, device=dev)
        ids_mask = torch.empty(sum(counts) - n_keep, dtype=torch.int64, device=dev)
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())          # CPU generator: follows torch.manual_seed
        L.run("mask_split", L.ptr(frame_starts), len(counts), float(keep_frac), seed, L.ptr(ids_keep), L.ptr(ids_mask),
              L.stream_ptr(dev))
        return ids_keep, ids_mask

    # ------------------------------------------------------------------ hot path
    def extract_feat(self, points, ids=None):
        batch_size = len(points)
        side = getattr(self, "scatter_stream", None)
        pb = scatter_frames(self.geom, points, side=side)
        fused = getattr(self, "fused_loss", True)
        layouts = None
        if side is not None and fused and getattr(self.backbone, "sra_impl", "tc3") != "glue":
            # everything that depends on the scatter result and the mask split only — the split itself, both window
            # layouts, the geometric targets — is queued on the input stream and runs under the VFE forward
            main = torch.cuda.current_stream(pb.points.device)
            v = pb.n_pillars
            with torch.cuda.stream(side), L.stream_override(side), torch.no_grad():
                ids_keep, ids_mask = ids if ids is not None else self.get_vanilla_mask_index(
                    pb.pillar_coors[:v], batch_size, pb.pillars_per_frame(), pb.counts[4:])
                layouts = self.backbone.build_layouts(pb, ids_keep, ids_mask)
                normal, curv = pb.geom_targets()
                ready = torch.cuda.Event()
                ready.record(side)
            voxel_features, feature_coors = self.voxel_encoder(pb)
            main.wait_event(ready)
            for t in (ids_keep, ids_mask, normal, curv):
                t.record_stream(main)
            for lay in layouts:
                lay.hand_over(main)
            low = low_mask = med = med_mask = top = None
            normal_m = normal
        else:
            voxel_features, feature_coors = self.voxel_encoder(pb)
            ids_keep, ids_mask = ids if ids is not None else self.get_vanilla_mask_index(
                feature_coors, batch_size, pb.pillars_per_frame(), pb.counts[4:])
            with torch.no_grad():
                normal, curv = pb.geom_targets()
                if fused:       # the fused loss reads the CSR sub-voxel lists directly: no dense targets at all
                    low = low_mask = med = med_mask = top = None
                    normal_m = normal
                else:
                    low, low_mask, med, med_mask, top = pb.dense_targets(ids_mask)
                    normal_m = normal.index_select(0, ids_mask)
        if getattr(self, "keep_targets", False):   # parity harness: expose exactly what this step regressed against
            self.last_targets = dict(pillar_batch=pb, normal=normal, curvature=curv, ids_keep=ids_keep,
                                     ids_mask=ids_mask)
        if layouts is not None:     # the layouts are built: the backbone needs no coordinates, only the pillar rows
            x = self.backbone(voxel_features.index_select(0, ids_keep), None, None, batch_size, pillar_batch=pb,
                              rows_keep=ids_keep, rows_mask=ids_mask, layouts=layouts)
        else:
            x = self.backbone(voxel_features.index_select(0, ids_keep), feature_coors.index_select(0, ids_keep),
                              feature_coors.index_select(0, ids_mask), batch_size, pillar_batch=pb, rows_keep=ids_keep,
                              rows_mask=ids_mask)
        if fused:
            return x, pb, ids_mask, normal_m
        return x, low, low_mask, med, med_mask, top, normal_m, None, None

    def forward_train(self, points, img_metas=None, gt_bboxes_3d=None, gt_labels_3d=None, gt_bboxes_ignore=None,
                      ids=None):
        for p in points:
            L.require_cuda(p, "points")
        out = self.extract_feat(points, ids)
        if len(out) == 4:
            x, pb, ids_mask, normal = out
            return self.fused_loss_from_csr(x, pb, ids_mask, normal)
        x, c_low, m_low, c_med, m_med, c_top, n_low, n_med, n_top = out
        reg_low, reg_med, reg_top, nor_low, nor_med, nor_top, cls_low, cls_med = x
        return self.forward_loss(c_low, m_low, c_med, m_med, c_top, n_low, n_med, n_top, reg_low, reg_med, reg_top,
                                 nor_low, nor_med, nor_top, cls_low, cls_med)

    LOSS_KEYS = ("loss_curv_around", "loss_centroid_low", "loss_centroid_med", "loss_centroid_top", "loss_cls_low",
                 "loss_cls_med")

    def fused_loss_from_csr(self, x, pb, ids_mask, normal):
        """forward_loss (…_ssl.py:837-902) as one fused kernel over the CSR targets (csrc/loss.cu)."""
        reg_low, reg_med, reg_top, nor_low, nor_med, nor_top, cls_low, cls_med = x
        nor_pred = nor_top if (nor_low is None and nor_med is None) else nor_low
        weights = (self.loss_ratio_low, self.loss_ratio_med, self.loss_ratio_top, self.loss_ratio_low_nor,
                   self.cls_loss_ratio_low, self.cls_loss_ratio_med)
        fused = getattr(self.backbone, "_fused_heads", None)
        if fused is not None:       # predictions are column slices of the fused head GEMMs: no copies
            out_c, out_d, fh = fused
            self.backbone._fused_heads = None
            losses = _GeomLossFusedFn.apply(out_c, out_d, fh, pb, ids_mask, normal, weights)
        else:
            losses = _GeomLossFn.apply(reg_low, reg_med, reg_top, nor_pred, cls_low, cls_med, pb, ids_mask, normal, weights)
        self.last_loss_vector = losses          # [6]; FlatTrainer sums this instead of six 0-dim views
        return {k: losses[i] for i, k in enumerate(self.LOSS_KEYS)}

    def forward(self, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(**kwargs)
        raise NotImplementedError("the SSL detector has no test path")

    @staticmethod
    def _mse(pred, target, weight):
        per_row = ((pred - target) ** 2).mean(dim=-1)
        return per_row.sum() / per_row.shape[0] * weight

    def forward_loss(self, centroid_low, centroid_low_mask, centroid_med, centroid_med_mask, centroid_high,
                     centroid_normal_low, centroid_normal_med, centroid_normal_high, reg_pred_low, reg_pred_med,
                     reg_pred_high, nor_pred_low, nor_pred_med, nor_pred_high, cls_pred_low=None, cls_pred_med=None):
        """…_ssl.py:837-902, mse_loss branch."""
        lm, mm = centroid_low_mask.reshape(-1), centroid_med_mask.reshape(-1)
        loss_low = self._mse(reg_pred_low.reshape(-1, 3)[lm], centroid_low.reshape(-1, 3)[lm], self.loss_ratio_low)
        loss_med = self._mse(reg_pred_med.reshape(-1, 3)[mm], centroid_med.reshape(-1, 3)[mm], self.loss_ratio_med)
        loss_top = self._mse(reg_pred_high, centroid_high, self.loss_ratio_top)
        nor_pred = nor_pred_high if (nor_pred_low is None and nor_pred_med is None) else nor_pred_low
        loss_nor = self._mse(nor_pred, centroid_normal_low, self.loss_ratio_low_nor)
        loss_cls_low = self.cls_loss(cls_pred_low.reshape(-1, 2), lm.long()) * self.cls_loss_ratio_low
        loss_cls_med = self.cls_loss(cls_pred_med.reshape(-1, 2), mm.long()) * self.cls_loss_ratio_med
        return dict(loss_curv_around=loss_nor, loss_centroid_low=loss_low, loss_centroid_med=loss_med,
                    loss_centroid_top=loss_top, loss_cls_low=loss_cls_low, loss_cls_med=loss_cls_med)
