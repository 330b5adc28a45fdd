"""NuScenesDatasetSSL + BatchLoader on synthetic files: the reference's pipeline list is taken
unchanged, the host stages reproduce data.py's (reference-pinned) loaders, sharding follows the restated
DistributedGroupSampler."""
import pickle

import numpy as np
import pytest

from geomae_b200.data import NuScenesSSLIndex
from geomae_b200.dataset import BatchLoader, NuScenesDatasetSSL, build_dataset, epoch_indices
from tests.test_data_step_cpu import _write_scene

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
# configs/mae_sst/…6x_1e-5.py:167-197, verbatim structure
TRAIN_PIPELINE = [
    dict(type="LoadPointsFromFile", coord_type="LIDAR", load_dim=5, use_dim=5, file_client_args=dict(backend="disk")),
    dict(type="LoadPointsFromMultiSweeps", sweeps_num=9, use_dim=[0, 1, 2, 3, 4], file_client_args=dict(backend="disk"),
         pad_empty_sweeps=True, remove_close=True),
    dict(type="GlobalRotScaleTrans", rot_range=[-0.3925, 0.3925], scale_ratio_range=[0.95, 1.05],
         translation_std=[0, 0, 0]),
    dict(type="RandomFlip3D", sync_2d=False, flip_ratio_bev_horizontal=0.5, flip_ratio_bev_vertical=0.5),
    dict(type="PointsRangeFilter", point_cloud_range=PC_RANGE),
    dict(type="PointShuffle"),
    dict(type="DefaultFormatBundle3D", class_names=["car"]),
    dict(type="Collect3D", keys=["points"]),
]


def write_dataset(tmp_path, n_scenes=5):
    rng = np.random.default_rng(0)
    infos = []
    for s in range(n_scenes):
        info, _ = _write_scene(tmp_path, rng, n_sweeps=(0, 3, 12)[s % 3], tag=f"s{s}")
        info.update(token=f"tok{s}", timestamp=1.6e15 + 1e6 * s)
        infos.append(info)
    ann = str(tmp_path / "nuscenes_ssl_infos_train.pkl")
    with open(ann, "wb") as f:
        pickle.dump(dict(infos=infos[::-1], metadata=dict(version="v1.0-trainval")), f)
    return ann


def test_dataset_takes_the_reference_pipeline_and_splits_host_from_device_stages(tmp_path):
    ann = write_dataset(tmp_path)
    ds = build_dataset(dict(type="NuScenesDatasetSSL", data_root=str(tmp_path), ann_file=ann, pipeline=TRAIN_PIPELINE,
                            classes=["car"], modality=dict(use_lidar=True), test_mode=False, box_type_3d="LiDAR"))
    assert isinstance(ds, NuScenesDatasetSSL) and len(ds) == 5 and ds.version == "v1.0-trainval"
    assert [ds.get_data_info(i)["sample_idx"] for i in range(5)] == [f"tok{i}" for i in range(5)]   # sorted by time
    assert ds.point_cloud_range == tuple(PC_RANGE) and (ds.flag == 0).all()
    index = NuScenesSSLIndex(ann)
    for i in range(5):
        item = ds.__getitem__(i, rng=np.random.RandomState(7))
        ref = index.load_frame(i, sweeps_num=9, test_mode=False, rng=np.random.RandomState(7))
        assert item["points"].dtype == np.float32 and np.array_equal(item["points"], ref)
        a = item["aug"]
        assert -0.3925 <= a.rotation <= 0.3925 and 0.95 <= a.scale <= 1.05
    with pytest.raises(NotImplementedError):
        NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE + [dict(type="ObjectSample")])
    plain = NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE[:2])        # test pipeline: no random stages
    a = plain[0]["aug"]
    assert (a.rotation, a.scale, a.flip_horizontal, a.flip_vertical) == (0.0, 1.0, False, False)


def test_epoch_indices_partition_and_reshuffle():
    n, spg, world = 23, 4, 2
    for epoch in (0, 1):
        shares = [epoch_indices(n, spg, r, world, epoch, seed=3) for r in range(world)]
        assert all(len(s) == 12 for s in shares)                         # ceil(23 / 8) * 8 / 2
        both = shares[0] + shares[1]
        assert set(both) == set(range(n)) and len(both) == 24            # every sample once, one wrapped around
    assert epoch_indices(n, spg, 0, world, 0, seed=3) != epoch_indices(n, spg, 0, world, 1, seed=3)
    assert epoch_indices(n, spg, 0, world, 5, seed=3) == epoch_indices(n, spg, 0, world, 5, seed=3)
    assert epoch_indices(8, 4, 1, 2, 0, shuffle=False) == [4, 5, 6, 7]


def test_batch_loader_is_deterministic_and_covers_the_epoch(tmp_path):
    ann = write_dataset(tmp_path, n_scenes=6)
    ds = NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE)
    seen = []
    for rank in range(2):
        loader = BatchLoader(ds, samples_per_gpu=2, rank=rank, world=2, seed=1, workers=3, prefetch=2, pin_memory=False)
        assert len(loader) == 2
        loader.set_epoch(4)
        first = [(tuple(p.shape[0] for p in pts), tuple(a.rotation for a in augs)) for pts, augs in loader]
        again = [(tuple(p.shape[0] for p in pts), tuple(a.rotation for a in augs)) for pts, augs in loader]
        assert first == again and len(first) == 2                        # thread timing does not change the stream
        for pts, augs in loader:
            assert len(pts) == 2 == len(augs) and all(p.dtype.is_floating_point and p.shape[1] == 5 for p in pts)
        seen += epoch_indices(len(ds), 2, rank, 2, 4, seed=1)
    assert sorted(seen) == [0, 0, 1, 2, 3, 4, 5, 5] or set(seen) == set(range(6))


def test_device_merge_mode_keeps_the_sweeps_raw_with_the_same_draws(tmp_path):
    from geomae_b200.data import RawSweeps
    ann = write_dataset(tmp_path)
    host = NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE)
    dev = NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE, device_merge=True)
    for i in range(len(host)):
        a = host.__getitem__(i, rng=np.random.RandomState(3))
        b = dev.__getitem__(i, rng=np.random.RandomState(3))
        raw = b["points"]
        assert isinstance(raw, RawSweeps) and raw.params.shape == (len(raw.arrays), 16)
        assert a["aug"] == b["aug"]                                   # the sweep choice consumed the same draws
        assert raw.n_points >= a["points"].shape[0]                   # close points still inside
        assert np.array_equal(raw.arrays[0][:, :4], a["points"][: raw.arrays[0].shape[0], :4])   # key frame first, untouched
        assert raw.params[0, 13] == -1.0 and (raw.params[1:, 13] == 1.0).all()
    loader = BatchLoader(dev, samples_per_gpu=2, workers=2, pin_memory=False)
    pts, augs = next(iter(loader))
    assert all(isinstance(p, RawSweeps) for p in pts) and len(augs) == 2
