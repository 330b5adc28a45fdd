"""Benchmark workloads = BASELINE.json `configs` as (frame generator arguments, model geometry, samples per GPU).

The reference's model code is generic in range / voxel sizes / grid — it reads them from the config
(configs/mae_sst/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5.py:14-24) — so the other shapes are the
same config with only those values replaced (`with_geometry`), exactly what the golden generator does to the
reference's own config for the parity cases `waymo_b2` / `dense_b1`."""
from __future__ import annotations

NUS_GEOMETRY = dict(pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), voxel_size=(0.256, 0.256, 8),
                    sub_voxel_size_med=(0.128, 0.128, 2), sub_voxel_size_low=(0.064, 0.064, 1), grid_size=(1, 400, 400))
# Waymo range / pillar size of configs/sst/sst_waymoD5_1x_3class_8heads.py:8-10, sub-voxels by the mae_sst ratios
WAYMO_GEOMETRY = dict(pc_range=(-74.88, -74.88, -2.0, 74.88, 74.88, 4.0), voxel_size=(0.32, 0.32, 6),
                      sub_voxel_size_med=(0.16, 0.16, 1.5), sub_voxel_size_low=(0.08, 0.08, 0.75), grid_size=(1, 468, 468))
DENSE_GEOMETRY = dict(pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), voxel_size=(0.1, 0.1, 8),
                      sub_voxel_size_med=(0.05, 0.05, 2), sub_voxel_size_low=(0.025, 0.025, 1), grid_size=(1, 1024, 1024))

WORKLOADS = {
    # BASELINE.json configs[1] / [2]: the configuration the metric is quoted on
    "nus": dict(label="mae_sst nuScenes config, synthetic 30k-pt sweeps, 1xB200 (BASELINE.json configs[1])",
                frame=dict(preset="nuscenes", sweeps=1), geometry=None, samples_per_gpu=4),
    # the real nuScenes input of the config: key frame + 9 sweeps (…6x_1e-5.py:176), ~279 k points / frame
    "nus10sweep": dict(label="mae_sst nuScenes config, synthetic 10-sweep frames (~280k pts)",
                       frame=dict(preset="nuscenes", sweeps=10), geometry=None, samples_per_gpu=4),
    # BASELINE.json configs[3]
    "waymo": dict(label="Waymo-shaped synthetic ~155k-pt frames, 0.32 m voxels (BASELINE.json configs[3])",
                  frame=dict(preset="waymo", sweeps=1), geometry=WAYMO_GEOMETRY, samples_per_gpu=4),
    # BASELINE.json configs[4]
    "dense": dict(label="dense-grid stress: 0.1 m voxels, ~200k non-empty pillars/frame (BASELINE.json configs[4])",
                  frame=dict(preset="nuscenes", sweeps=10, point_scale=2.4), geometry=DENSE_GEOMETRY, samples_per_gpu=2),
}


def with_geometry(model_cfg, geometry, base=NUS_GEOMETRY):
    """A copy of a mae_sst model config with range, the three voxel sizes (recognised by value: the config repeats them
    in seven sub-dicts), grid_size, spatial_shape and output_shape replaced."""
    if geometry is None:
        return model_cfg
    swap = {tuple(base[k]): tuple(geometry[k]) for k in ("voxel_size", "sub_voxel_size_med", "sub_voxel_size_low")}
    gz, gy, gx = geometry["grid_size"]

    def walk(node):
        if isinstance(node, dict):
            out = type(node)()
            for k, v in node.items():
                if k == "point_cloud_range":
                    out[k] = list(geometry["pc_range"])
                elif isinstance(k, str) and k.startswith(("voxel_size", "sub_voxel_size")) and isinstance(v, (list, tuple)) and tuple(v) in swap:
                    out[k] = swap[tuple(v)]
                elif k == "grid_size":
                    out[k] = (gz, gy, gx)
                elif k == "spatial_shape":
                    out[k] = [gz, gy, gx]
                elif k == "output_shape":
                    out[k] = [gy, gx]
                else:
                    out[k] = walk(v)
            return out
        if isinstance(node, (list, tuple)):
            return type(node)(walk(v) for v in node)
        return node
    return walk(model_cfg)
