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


def test_train_step_from_host_with_device_augmentation_matches_preaugmented_frames():
    """FlatTrainer.train_step_from_host(augs=...) == the same step on frames augmented + filtered by the oracle on the
    host (same mask split, same weights): the device data step changes where the work happens, not the result."""
    import os
    import geomae_b200 as G
    from geomae_b200.data import draw_augmentation, frame_params
    from geomae_b200.registry import Config
    from geomae_b200.synthetic import make_frame
    from geomae_b200.train import FlatTrainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs/mae_sst/geomae_nus_pretrain.py"))
    frames = [make_frame(91, point_scale=0.5), make_frame(92, point_scale=0.3)]
    rs = np.random.RandomState(5)
    augs = [draw_augmentation(rs) for _ in frames]
    pre = O.augment_filter(frames, frame_params(augs).numpy(), O.PathConfig().pc_range)
    assert sum(f.shape[0] for f in pre) < sum(f.shape[0] for f in frames)

    def run(feed):
        torch.manual_seed(0)
        model = G.build_detector(cfg.model).to(DEV)
        model.set_impl("tc3")
        model.train()
        tr = FlatTrainer(model, lr=1e-4)
        torch.manual_seed(3)                        # the mask split draws its seed from the CPU generator
        loss, parts = feed(tr)
        return {k: float(v) for k, v in parts.items()}, tr.flat_param.clone()

    host = [torch.from_numpy(f).pin_memory() for f in frames]
    la, pa = run(lambda tr: tr.train_step_from_host(host, augs=augs))
    lb, pb = run(lambda tr: tr.train_step([torch.from_numpy(f).to(DEV) for f in pre]))
    from tests.golden_util import assert_same_step
    assert_same_step(la, lb, tol=1e-5)           # same points in the same order
    # one AdamW step moves every weight by about lr; a gradient that is pure rounding noise may flip its direction
    assert (pa - pb).abs().max().item() <= 2.01e-4 and (pa - pb).abs().mean().item() <= 1e-7


def _scene(tmp_path, n_sweeps, seed):
    from tests.test_data_step_cpu import _write_scene
    rng = np.random.default_rng(seed)
    info, _ = _write_scene(tmp_path, rng, n_sweeps, tag=f"g{seed}")
    return dict(pts_filename=info["lidar_path"], sweeps=info["sweeps"], timestamp=info["timestamp"] / 1e6)


def test_sweep_merge_matches_host_loader(tmp_path):
    """geomae_sweep_merge vs data.load_multi_sweeps (bit-exact vs the reference's LoadPointsFromMultiSweeps on CPU):
    same sweeps, same survivors in the same order, same time channel; coordinates equal except where the BLAS dot
    product of the host (fused multiply-adds) and the kernel's separately rounded float64 products round a float32
    tie differently (<= 1 ulp, a vanishing fraction)."""
    from geomae_b200.data import load_multi_sweeps, load_multi_sweeps_raw, read_points_bin, sweep_merge
    samples, refs = [], []
    for n_sweeps, seed in ((0, 1), (4, 2), (14, 3)):
        info = _scene(tmp_path, n_sweeps, seed)
        key = read_points_bin(info["pts_filename"], 5, 5)
        refs.append(load_multi_sweeps(key, info, rng=np.random.RandomState(9)))
        samples.append(load_multi_sweeps_raw(key, info, rng=np.random.RandomState(9)))
    segs = [a for s in samples for a in s.arrays]
    offs = np.concatenate([[0], np.cumsum([a.shape[0] for a in segs])]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(segs)).to(DEV)
    par = torch.from_numpy(np.concatenate([s.params for s in samples])).to(DEV)
    out, out_off = sweep_merge(pts, torch.from_numpy(offs).to(DEV), par)
    o = out_off.cpu().numpy()
    first = np.concatenate([[0], np.cumsum([len(s.arrays) for s in samples])])
    got_all = out.cpu().numpy()
    for i, ref in enumerate(refs):
        got = got_all[o[first[i]]:o[first[i + 1]]]
        assert got.shape == ref.shape
        assert np.array_equal(got[:, 3:], ref[:, 3:])                      # intensity, time channel: exact
        diff = got[:, :3] != ref[:, :3]
        assert diff.mean() < 1e-4
        ulp = np.spacing(np.abs(ref[:, :3]).astype(np.float32))
        assert (np.abs(got[:, :3] - ref[:, :3]) <= ulp).all()
    assert o[-1] < pts.shape[0]                                            # close points were dropped


def test_train_step_from_host_merges_raw_sweeps_on_the_device(tmp_path):
    """RawSweeps samples through train_step_from_host (sweep merge -> augmentation -> scatter, all on the input
    stream) == the same step fed with host-merged frames."""
    import os
    import geomae_b200 as G
    from geomae_b200.data import draw_augmentation, load_multi_sweeps, load_multi_sweeps_raw, read_points_bin
    from geomae_b200.registry import Config
    from geomae_b200.train import FlatTrainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs/mae_sst/geomae_nus_pretrain.py"))
    raw, merged = [], []
    for n_sweeps, seed in ((3, 5), (0, 6)):
        info = _scene(tmp_path, n_sweeps, seed)
        key = read_points_bin(info["pts_filename"], 5, 5)
        merged.append(torch.from_numpy(load_multi_sweeps(key, info, rng=np.random.RandomState(4))).pin_memory())
        raw.append(load_multi_sweeps_raw(key, info, rng=np.random.RandomState(4)))
    rs = np.random.RandomState(8)
    augs = [draw_augmentation(rs) for _ in raw]

    def run(feed):
        torch.manual_seed(0)
        model = G.build_detector(cfg.model).to(DEV)
        model.set_impl("tc3")
        model.train()
        tr = FlatTrainer(model, lr=1e-4)
        torch.manual_seed(3)
        return {k: float(v) for k, v in feed(tr)[1].items()}

    for a in (None, augs):
        la = run(lambda tr: tr.train_step_from_host(raw, augs=a))
        lb = run(lambda tr: tr.train_step_from_host(merged, augs=a))
        from tests.golden_util import assert_same_step
        assert_same_step(la, lb)


def test_ragged_batch_through_the_trainer_paths():
    """An empty sample and a 3-point sample next to a normal one through every trainer entry (resident, host-fed,
    host-fed with device augmentation): the input stream's offsets / totals and the per-frame mask split cope."""
    import os
    import geomae_b200 as G
    from geomae_b200.data import draw_augmentation
    from geomae_b200.registry import Config
    from geomae_b200.synthetic import make_frame
    from geomae_b200.train import FlatTrainer
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs/mae_sst/geomae_nus_pretrain.py"))
    frames = [make_frame(71, point_scale=0.1), np.zeros((0, 5), np.float32), make_frame(72, point_scale=0.05)[:3].copy()]
    torch.manual_seed(0)
    model = G.build_detector(cfg.model).to(DEV).train()
    model.set_impl("tc1")
    tr = FlatTrainer(model, lr=1e-4)
    rs = np.random.RandomState(2)
    host = [torch.from_numpy(f).pin_memory() for f in frames]
    losses = [tr.train_step([torch.from_numpy(f).to(DEV) for f in frames])[0],
              tr.train_step_from_host(host)[0],
              tr.train_step_from_host(host, augs=[draw_augmentation(rs) for _ in frames])[0]]
    vals = [float(v) for v in losses]
    assert all(np.isfinite(vals)) and all(v > 0 for v in vals), vals
