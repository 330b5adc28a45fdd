"""GPU parity of the fused augmentation + range filter + compaction (geomae_augment_filter) against
the oracle's restatement — bit exact: same fp32 operations, each rounded separately — and composed with the voxel
scatter."""
import numpy as np
import pytest
import torch

from oracle import geomae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def batch(frames):
    pts = torch.from_numpy(np.concatenate(frames, axis=0)).to(DEV)
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int32)
    return pts, torch.from_numpy(offs).to(DEV)


def test_augment_filter_matches_oracle_bit_exact():
    from geomae_b200.data import Augmentation, augment_filter, frame_params
    from geomae_b200.synthetic import make_frame
    cfg = O.PathConfig()
    frames = [make_frame(81), np.zeros((0, 5), np.float32), make_frame(82, sweeps=2), make_frame(83)[:3],
              make_frame(84, point_scale=0.3), np.zeros((0, 5), np.float32)]
    augs = [Augmentation(0.39, 1.05, False, False), Augmentation(0.1, 1.0, True, True),
            Augmentation(-0.2, 0.95, True, False), Augmentation(0.0, 1.0, False, True),
            Augmentation(-0.3925, 1.02, True, True), Augmentation()]
    pts, offs = batch(frames)
    out, out_off = augment_filter(pts, offs, augs, cfg.pc_range)
    ref = O.augment_filter(frames, frame_params(augs).numpy(), cfg.pc_range)
    off = out_off.cpu().numpy()
    assert off[0] == 0 and list(np.diff(off)) == [r.shape[0] for r in ref]
    got = out.cpu().numpy()
    for b, r in enumerate(ref):
        assert np.array_equal(got[off[b]:off[b + 1]], r), b
    assert 0 < off[-1] < pts.shape[0]                      # the rotated corners of the range fall outside


def test_filtered_batch_feeds_the_scatter():
    from geomae_b200.data import Augmentation, augment_filter, frame_params
    from geomae_b200.synthetic import make_frame
    from geomae_b200.voxel import VoxelGeometry, scatter_frames
    cfg = O.PathConfig()
    frames = [make_frame(85), make_frame(86)]
    augs = [Augmentation(0.2, 0.97, True, False), Augmentation(-0.35, 1.03, False, True)]
    pts, offs = batch(frames)
    out, out_off = augment_filter(pts, offs, augs, cfg.pc_range)
    off = out_off.tolist()
    geom = VoxelGeometry(cfg.pc_range, cfg.voxel_size, cfg.sub_voxel_size_med, cfg.sub_voxel_size_low,
                         cfg.sub_voxel_ratio_med, cfg.sub_voxel_ratio_low)
    pb = scatter_frames(geom, [out[off[b]:off[b + 1]] for b in range(2)])
    ref_frames = O.augment_filter(frames, frame_params(augs).numpy(), cfg.pc_range)
    rows, inv, cnt = O.unique_rows(O.batch_voxelize(ref_frames, cfg.voxel_size, cfg.pc_range))
    v = pb.n_pillars
    assert np.array_equal(pb.pillar_coors[:v].cpu().numpy(), rows)
    assert np.array_equal(pb.point_pillar[:off[-1]].cpu().numpy(), inv)


def test_empty_batch():
    from geomae_b200.data import Augmentation, augment_filter
    cfg = O.PathConfig()
    pts = torch.zeros((0, 5), device=DEV)
    offs = torch.zeros(3, dtype=torch.int32, device=DEV)
    out, out_off = augment_filter(pts, offs, [Augmentation(), Augmentation()], cfg.pc_range)
    assert out.shape == (0, 5) and out_off.tolist() == [0, 0, 0]
