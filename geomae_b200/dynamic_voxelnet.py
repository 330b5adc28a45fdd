"""DynamicVoxelNet — the fine-tune consumer of a GeoMAE checkpoint: mirror of
mmdet3d/models/detectors/dynamic_voxelnet.py:10-83 up to and including ``extract_feat``
(voxelize -> DynamicScatterVFE -> SSTInputLayer -> SSTSecondPretrainedv1 [-> neck]).

The detection head (Anchor3DHead in the pre_sst config), its losses and the SECONDFPN neck are outside SURVEY §8; a config that names them
is accepted (so the reference's config file builds), the sub-configs are kept on the module, and ``forward_train`` /
``simple_test`` say so loudly instead of pretending.  ``load_pretrained`` is what ``load_from = …/epoch_72.pth`` does
in the reference (configs/pre_sst/…6x_1e-5.py:280; mmcv load_checkpoint, strict=False): tensors whose keys and shapes
match are copied, everything else is reported."""
from __future__ import annotations

import torch
from torch import nn

from .registry import DETECTORS, MIDDLE_ENCODERS, build_backbone, build_voxel_encoder
from .voxel import VoxelGeometry, Voxelization, scatter_frames


@DETECTORS.register_module()
class DynamicVoxelNet(nn.Module):
    def __init__(self, voxel_layer, voxel_encoder, middle_encoder, backbone, centerpoint_head=False, neck=None,
                 bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        self.voxel_layer = Voxelization(**voxel_layer)
        self.voxel_encoder = build_voxel_encoder(voxel_encoder)
        self.middle_encoder = MIDDLE_ENCODERS.build(middle_encoder)
        self.backbone = build_backbone(backbone)
        self.neck_cfg, self.bbox_head_cfg = neck, bbox_head      # not built: outside SURVEY §8
        self.centerpoint_head, self.train_cfg, self.test_cfg = centerpoint_head, train_cfg, test_cfg
        vs, rng = tuple(voxel_layer["voxel_size"]), tuple(voxel_layer["point_cloud_range"])
        self.geom = VoxelGeometry(rng, vs, vs, vs, (1, 1, 1), (1, 1, 1))

    with_neck = False

    def voxelize(self, points):
        """:55-77 — concatenated points and (b, z, y, x) coordinates."""
        coors = [nn.functional.pad(self.voxel_layer(p), (1, 0), value=i) for i, p in enumerate(points)]
        return torch.cat(points, dim=0), torch.cat(coors, dim=0)

    def extract_feat(self, points, img_metas=None):
        """:38-53.  One scatter pass feeds the VFE; the window drop and the CSR layout run on the pillar coordinates."""
        batch_size = len(points)
        pb = scatter_frames(self.geom, points)
        voxel_features, feature_coors = self.voxel_encoder(pb)
        x = self.middle_encoder(voxel_features, feature_coors, batch_size)
        return self.backbone(x)

    def forward_train(self, *args, **kwargs):
        raise NotImplementedError("DynamicVoxelNet here is the feature extractor of the fine-tune consumer; the "
                                  "detection head and its losses are outside the GeoMAE pre-training path (SURVEY §8)")

    simple_test = aug_test = forward_train

    def set_impl(self, impl: str):
        self.backbone.set_sra_impl(impl)
        self.voxel_encoder.tc_precision = {"tc3": 3, "tc1": 1}[impl]
        return self

    def load_pretrained(self, checkpoint):
        """checkpoint: path of a ``.pth`` written by the pre-training run, or its dict / bare state_dict.
        -> (loaded keys, keys of this model left untouched, checkpoint keys with no destination)."""
        if isinstance(checkpoint, str):
            checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
        sd = checkpoint.get("state_dict", checkpoint)
        own = self.state_dict()
        loaded, unexpected = [], []
        with torch.no_grad():
            for k, v in sd.items():
                k = k[7:] if k.startswith("module.") else k
                if k in own and own[k].shape == v.shape:
                    own[k].copy_(v)
                    loaded.append(k)
                else:
                    unexpected.append(k)
        done = set(loaded)
        return loaded, [k for k in own if k not in done], unexpected
