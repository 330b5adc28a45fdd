"""geomae_b200 — B200-native implementation of GeoMAE's masked-pretraining hot path.

Importing the package registers the reference's registry names (DETECTORS
'MultiSubVoxelDynamicVoxelNetSSL', BACKBONES 'MultiMAESSTSPChoose', VOXEL_ENCODERS
'DynamicScatterVFE', NORM_LAYERS 'naiveSyncBN1d', LOSSES 'SmoothL1Loss'/'CrossEntropyLoss'; and for the fine-tune
consumer DETECTORS 'DynamicVoxelNet', MIDDLE_ENCODERS 'SSTInputLayer', BACKBONES 'SSTSecondPretrainedv1',
NORM_LAYERS 'naiveSyncBN2d')."""
from . import backbone, detector, dynamic_voxelnet, losses, norm, sst_input_layer, sst_second, voxel_encoder  # noqa: F401  (registration side effects)
from .registry import Config, build_detector, build_model  # noqa: F401
from .voxel import Voxelization, VoxelGeometry, scatter_frames  # noqa: F401
from . import ops  # noqa: F401  (the reference's `mmdet3d.ops` names for this path)
