# GeoMAE masked pre-training, nuScenes geometry — the same hyper-parameters as the reference's
# configs/mae_sst/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5.py (model part only;
# that file itself also loads unchanged through geomae_b200.Config when the reference tree is present,
# see tests/test_config_cpu.py).  Written out independently so tests/bench run without the reference.
pc_range = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
pillar = (0.256, 0.256, 8)
sub_med, sub_low = (0.128, 0.128, 2), (0.064, 0.064, 1)
ratio_med, ratio_low = (4, 2, 2), (8, 4, 4)          # z, y, x sub-voxels per pillar
win = (12, 12)
buckets_train = {0: dict(max_tokens=56, drop_range=(0, 56)), 1: dict(max_tokens=144, drop_range=(56, 100000))}
buckets_test = {0: dict(max_tokens=32, drop_range=(0, 32)), 1: dict(max_tokens=72, drop_range=(32, 72)),
                2: dict(max_tokens=144, drop_range=(72, 100000))}


def _dyn(size):
    return dict(voxel_size=size, max_num_points=-1, point_cloud_range=pc_range, max_voxels=(-1, -1))


def _hard(size, pts, vox):
    return dict(voxel_size=size, max_num_points=pts, point_cloud_range=pc_range, max_voxels=(vox, vox))


model = dict(
    type='MultiSubVoxelDynamicVoxelNetSSL',
    normalize_sub_voxel=True, mse_loss=True, cls_sub_voxel=True,
    loss=dict(type='SmoothL1Loss', reduction='mean', loss_weight=1.0),
    spatial_shape=[1, 400, 400], grid_size=(1, 400, 400),
    loss_ratio_low=10.0, loss_ratio_med=8.0, loss_ratio_top=10.0,
    loss_ratio_low_nor=4.0, loss_ratio_med_nor=0, loss_ratio_top_nor=0,
    cls_loss_ratio_low=5.0, cls_loss_ratio_med=2.0,
    random_mask_ratio=0.7,
    sub_voxel_ratio_low=ratio_low, sub_voxel_ratio_med=ratio_med,
    voxel_layer=_dyn(pillar), sub_voxel_layer_low=_dyn(sub_low), sub_voxel_layer_med=_dyn(sub_med),
    hard_sub_voxel_layer_low=_hard(sub_low, 30, 140000),
    hard_sub_voxel_layer_med=_hard(sub_med, 50, 80000),
    hard_sub_voxel_layer_top=_hard(pillar, 100, 40000),
    voxel_encoder=dict(
        type='DynamicScatterVFE', in_channels=5, feat_channels=[64, 128], with_distance=False,
        with_cluster_center=True, with_voxel_center=True, voxel_size=pillar, point_cloud_range=pc_range,
        norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01)),
    backbone=dict(
        type='MultiMAESSTSPChoose', cls_sub_voxel=True, window_shape=win,
        shifts_list=[(0, 0), (win[0] // 2, win[1] // 2)], point_cloud_range=pc_range, voxel_size=pillar,
        shuffle_voxels=False, low=False, med=False, top=True,
        d_model=[128] * 6, nhead=[8] * 6, dim_feedforward=[256] * 6,
        sub_voxel_ratio_low=ratio_low, sub_voxel_ratio_med=ratio_med,
        encoder_num_blocks=6, decoder_num_blocks=2, output_shape=[400, 400],
        debug=True, drop_info=(buckets_train, buckets_test), pos_temperature=10000, normalize_pos=False),
)
data = dict(samples_per_gpu=4, workers_per_gpu=4)
# configs/_base_/schedules/cosine_2x.py:1-9 of the reference
optimizer = dict(type='AdamW', lr=1e-5, betas=(0.9, 0.999), weight_decay=0.05,
                 paramwise_cfg=dict(custom_keys={'norm': dict(decay_mult=0.)}))
optimizer_config = dict(grad_clip=dict(max_norm=10, norm_type=2))
# data / schedule of the reference config (…6x_1e-5.py:164-197,257-291); `tools/train.py` reads these
dataset_type, data_root = 'NuScenesDatasetSSL', 'data/nuscenes/'
class_names = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle', 'bicycle',
               'pedestrian', 'traffic_cone']
train_pipeline = [
    dict(type='LoadPointsFromFile', coord_type='LIDAR', load_dim=5, use_dim=5),
    dict(type='LoadPointsFromMultiSweeps', sweeps_num=9, use_dim=[0, 1, 2, 3, 4], pad_empty_sweeps=True,
         remove_close=True),
    dict(type='GlobalRotScaleTrans', rot_range=[-0.3925, 0.3925], scale_ratio_range=[0.95, 1.05],
         translation_std=[0, 0, 0]),
    dict(type='RandomFlip3D', sync_2d=False, flip_ratio_bev_horizontal=0.5, flip_ratio_bev_vertical=0.5),
    dict(type='PointsRangeFilter', point_cloud_range=pc_range),
    dict(type='PointShuffle'),
    dict(type='DefaultFormatBundle3D', class_names=class_names),
    dict(type='Collect3D', keys=['points']),
]
data.update(train=dict(type=dataset_type, data_root=data_root, ann_file=data_root + 'nuscenes_ssl_infos_train.pkl',
                       pipeline=train_pipeline, classes=class_names, test_mode=False, box_type_3d='LiDAR',
                       device_merge=True))
runner = dict(type='EpochBasedRunner', max_epochs=72)
