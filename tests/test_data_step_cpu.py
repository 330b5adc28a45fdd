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
