# Fine-tune consumer of a GeoMAE checkpoint, feature-extractor part — the same hyper-parameters
# as the reference's configs/pre_sst/m_sst_nus_second_pointpillar_fpn355_222_curv_07_ssl_data_wo_dbsampler_6x_1e-5.py
# :15-34,73-125 (voxelise -> DynamicScatterVFE -> SSTInputLayer -> SSTSecondPretrainedv1).  That file itself builds
# unchanged through geomae_b200.Config when the reference tree is present (tests/test_config_cpu.py); its neck and
# bbox_head sub-configs are kept on the detector but not built (outside §8).
pc_range = [-50, -50, -5.0, 50, 50, 3.0]
pillar = (0.25, 0.25, 8)
win = (12, 12)
shifts = [(0, 0), (win[0] // 2, win[1] // 2)]
buckets = {0: dict(max_tokens=32, drop_range=(0, 32)), 1: dict(max_tokens=72, drop_range=(32, 72)),
           2: dict(max_tokens=144, drop_range=(72, 1000))}
drop_info = (buckets, buckets)

model = dict(
    type='DynamicVoxelNet', centerpoint_head=False,
    voxel_layer=dict(voxel_size=pillar, max_num_points=-1, point_cloud_range=pc_range, max_voxels=(-1, -1)),
    voxel_encoder=dict(
        type='DynamicScatterVFE', in_channels=5, feat_channels=[64, 128], with_distance=False, voxel_size=pillar,
        with_cluster_center=True, with_voxel_center=True, point_cloud_range=pc_range,
        norm_cfg=dict(type='naiveSyncBN1d', eps=1e-3, momentum=0.01)),
    middle_encoder=dict(
        type='SSTInputLayer', window_shape=win, shifts_list=shifts, point_cloud_range=pc_range, voxel_size=pillar,
        shuffle_voxels=True, debug=True, drop_info=drop_info),
    backbone=dict(
        type='SSTSecondPretrainedv1', eval_flag=False, model_path='', d_model=[128] * 6, nhead=[8] * 6, num_blocks=6,
        dim_feedforward=[256] * 6, output_shape=[400, 400], conv_in_channels=128, conv_out_channels=[128, 128, 256],
        layer_nums=[3, 5, 5], layer_strides=[2, 2, 2], debug=True, drop_info=drop_info, pos_temperature=10000,
        normalize_pos=False, window_shape=win),
)
load_from = 'work_dirs/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5/epoch_72.pth'
