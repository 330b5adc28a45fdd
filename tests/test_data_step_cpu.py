"""The oracle's restatement of the pretraining pipeline's augmentation + range filter (oracle.augment_filter) against the
reference's own point class (mmdet3d/core/points/{base,lidar}_points.py, loaded by path: pure torch) driven the way
GlobalRotScaleTrans / RandomFlip3D / PointsRangeFilter drive it (datasets/pipelines/transforms_3d.py:670-768,125-160,
849-883)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from geomae_b200.data import Augmentation, draw_augmentation, frame_params
from geomae_b200.synthetic import make_frame
from oracle import geomae_oracle as O

REF_POINTS = "/root/reference/mmdet3d/core/points"


def reference_lidar_points():
    pkg = types.ModuleType("_ref_points")
    pkg.__path__ = [REF_POINTS]
    sys.modules["_ref_points"] = pkg
    mods = {}
    for name in ("base_points", "lidar_points"):
        spec = importlib.util.spec_from_file_location(f"_ref_points.{name}", os.path.join(REF_POINTS, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["lidar_points"].LiDARPoints


def reference_pipeline(LiDARPoints, frame, aug, pc_range):
    pts = LiDARPoints(torch.from_numpy(frame.copy()), points_dim=frame.shape[1])
    pts.rotate(aug.rotation)                                   # GlobalRotScaleTrans._rot_bbox_points (:679-685)
    pts.scale(aug.scale)                                       # ._scale_bbox_points (:709-711)
    pts.translate(np.zeros(3, np.float32))                     # ._trans_bbox_points with translation_std 0 (:662-665)
    if aug.flip_horizontal:
        pts.flip("horizontal")                                 # RandomFlip3D (:154-159)
    if aug.flip_vertical:
        pts.flip("vertical")
    mask = pts.in_range_3d(np.array(pc_range, np.float32))     # PointsRangeFilter (:868-870)
    return pts.tensor.numpy(), mask.numpy()


@pytest.mark.skipif(not os.path.exists(REF_POINTS), reason="reference tree not present")
def test_oracle_matches_reference_point_ops():
    LiDARPoints = reference_lidar_points()
    cfg = O.PathConfig()
    rng = np.random.RandomState(3)
    frames = [make_frame(71), make_frame(72, sweeps=2), make_frame(73, point_scale=0.2)]
    augs = [draw_augmentation(rng) for _ in frames]
    augs.append(Augmentation(0.3, 1.04, True, True))
    frames.append(make_frame(74))
    params = frame_params(augs).numpy()
    got = O.augment_filter(frames, params, cfg.pc_range)
    lo, hi = np.array(cfg.pc_range[:3], np.float32), np.array(cfg.pc_range[3:], np.float32)
    for f, a, p, g in zip(frames, augs, params, got):
        ref_xyz, ref_mask = reference_pipeline(LiDARPoints, f, a, cfg.pc_range)
        # torch's matmul may contract the two products of a row into an FMA: allow the last bit
        margin = np.minimum(np.abs(ref_xyz[:, :3] - lo), np.abs(ref_xyz[:, :3] - hi)).min(axis=1)
        clear = margin > 1e-4                                  # points not within 0.1 mm of a range face
        assert clear.mean() > 0.999
        mine_all = O.augment_filter([f], p[None], [-1e9, -1e9, -1e9, 1e9, 1e9, 1e9])[0]
        np.testing.assert_allclose(mine_all[:, :3], ref_xyz[:, :3], rtol=3e-7, atol=2e-6)
        assert np.array_equal(mine_all[:, 3:], ref_xyz[:, 3:])
        keep = ((mine_all[:, :3] > lo) & (mine_all[:, :3] < hi)).all(axis=1)
        assert np.array_equal(keep[clear], ref_mask[clear])
        assert g.shape[0] == int(keep.sum()) and np.array_equal(g, mine_all[keep])
        assert 0.3 < ref_mask.mean() <= 1.0


def test_draws_follow_the_configured_ranges():
    rng = np.random.RandomState(0)
    augs = [draw_augmentation(rng) for _ in range(2000)]
    rot = np.array([a.rotation for a in augs])
    sc = np.array([a.scale for a in augs])
    assert rot.min() >= -0.3925 and rot.max() <= 0.3925 and abs(rot.mean()) < 0.02
    assert sc.min() >= 0.95 and sc.max() <= 1.05
    assert 0.45 < np.mean([a.flip_horizontal for a in augs]) < 0.55
    assert 0.45 < np.mean([a.flip_vertical for a in augs]) < 0.55
    p = frame_params([Augmentation(0.25, 1.01, True, False), Augmentation(-0.1, 0.97, True, True)]).numpy()
    assert p.dtype == np.float32 and p.shape == (2, 4) and list(p[:, 3]) == [1.0, 3.0]
    np.testing.assert_allclose(p[0, :2] ** 2 @ np.ones(2), 1.0, atol=1e-6)


def _write_scene(tmp_path, rng, n_sweeps, tag="a"):
    """A synthetic key frame + sweeps on disk and the matching infos entry (nuscenes_ssl_converter layout)."""
    def bin_file(name, n):
        pts = np.concatenate([rng.uniform(-40, 40, (n, 2)), rng.uniform(-4, 2, (n, 1)), rng.uniform(0, 255, (n, 1)),
                              rng.integers(0, 32, (n, 1))], axis=1).astype(np.float32)
        pts[: n // 20, :2] *= 0.01                      # some points inside the 1 m "close" box
        path = str(tmp_path / name)
        pts.tofile(path)
        return path, pts
    key_path, key = bin_file(f"{tag}_key.pcd.bin", 3000)
    sweeps = []
    for i in range(n_sweeps):
        path, _ = bin_file(f"{tag}_sweep{i}.pcd.bin", 2500 + 10 * i)
        ang = rng.uniform(-0.05, 0.05)
        rot = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
        sweeps.append(dict(data_path=path, timestamp=1.6e15 - 5e4 * (i + 1), sensor2lidar_rotation=rot,
                           sensor2lidar_translation=rng.uniform(-0.5, 0.5, 3)))
    info = dict(token=f"tok{n_sweeps}", lidar_path=key_path, sweeps=sweeps, timestamp=1.6e15)
    return info, key


@pytest.mark.skipif(not os.path.exists(REF_POINTS), reason="reference tree not present")
@pytest.mark.parametrize("n_sweeps,test_mode", [(0, False), (4, False), (14, False), (14, True)])
def test_file_and_multi_sweep_loading_matches_reference(tmp_path, n_sweeps, test_mode):
    """read_points_bin + load_multi_sweeps against the reference's LoadPointsFromFile + LoadPointsFromMultiSweeps
    (datasets/pipelines/loading.py:337-443,100-235) configured as in …6x_1e-5.py:168-180, on files written here;
    NuScenesSSLIndex against NuScenesDatasetSSL.load_annotations / get_data_info (nuscenes_ssl_dataset.py:176-235)."""
    import pickle
    from geomae_b200.data import NuScenesSSLIndex, load_multi_sweeps, read_points_bin
    from oracle import ref_harness as H
    loading = H.load_pipeline_loading()
    rng = np.random.default_rng(n_sweeps)
    info, key = _write_scene(tmp_path, rng, n_sweeps)
    other, _ = _write_scene(tmp_path, rng, 0, tag="b")
    other.update(token="earlier", timestamp=1.5e15)
    ann = str(tmp_path / "nuscenes_ssl_infos_train.pkl")
    with open(ann, "wb") as f:
        pickle.dump(dict(infos=[info, other], metadata=dict(version="v1.0-trainval")), f)
    index = NuScenesSSLIndex(ann)
    assert len(index) == 2 and index.version == "v1.0-trainval"
    assert index.get_data_info(0)["sample_idx"] == "earlier"           # sorted by timestamp
    entry = index.get_data_info(1)
    assert entry["pts_filename"] == info["lidar_path"] and entry["timestamp"] == info["timestamp"] / 1e6

    load_file = loading.LoadPointsFromFile(coord_type="LIDAR", load_dim=5, use_dim=5)
    load_sweeps = loading.LoadPointsFromMultiSweeps(sweeps_num=9, use_dim=[0, 1, 2, 3, 4], pad_empty_sweeps=True,
                                                    remove_close=True, test_mode=test_mode)
    np.random.seed(5)
    results = load_sweeps(load_file(dict(pts_filename=entry["pts_filename"], sweeps=entry["sweeps"],
                                         timestamp=entry["timestamp"])))
    ref = results["points"].tensor.numpy()

    mine_key = read_points_bin(entry["pts_filename"], 5, 5)
    assert np.array_equal(mine_key, key)
    np.random.seed(5)
    mine = load_multi_sweeps(mine_key, entry, sweeps_num=9, test_mode=test_mode, rng=np.random)
    assert mine.dtype == np.float32 and mine.shape == ref.shape
    assert np.array_equal(mine, ref)
    np.random.seed(5)
    assert np.array_equal(index.load_frame(1, test_mode=test_mode, rng=np.random), ref)
    assert (mine[: key.shape[0], 4] == 0).all()
    if n_sweeps:
        assert mine.shape[0] > key.shape[0] and (mine[key.shape[0]:, 4] > 0).all()
