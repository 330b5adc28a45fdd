"""The CPU restatement (oracle/) must reproduce vectors captured from the unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import geomae_oracle as O
from tests.golden_util import align_sign, load_case


@pytest.fixture(scope="module")
def small():
    case, cfg, frames, g = load_case("small_b2")
    tgt = O.geometric_targets(frames, cfg, g["ids_mask"])
    return case, cfg, frames, g, tgt


def test_voxel_coords_bit_exact(small):
    _, _, frames, g, tgt = small
    for i, f in enumerate(frames):
        assert np.array_equal(f, g[f"points{i}"])
    for k in ("coors_top", "coors_med", "coors_low"):
        assert np.array_equal(tgt[k], g[k]), k


def test_sorted_pillars_and_centroids(small):
    _, _, _, g, tgt = small
    assert np.array_equal(tgt["pillar_coors"], g["pillar_coors"])
    assert np.array_equal(tgt["rows_med"], g["rows_med"])
    assert np.array_equal(tgt["rows_low"], g["rows_low"])
    assert np.array_equal(tgt["pillar_count"], g["count_top"])
    for k in ("top", "med", "low"):
        np.testing.assert_allclose(tgt[f"centroid_{k}"], g[f"centroid_{k}"], rtol=0, atol=2e-6)


def test_mask_split_matches_reference_generator(small):
    case, cfg, frames, g, tgt = small
    keep, mask = O.vanilla_mask_ids(tgt["pillar_coors"], len(frames), cfg.mask_ratio, case["mask_seed"])
    assert np.array_equal(keep, g["ids_keep"]) and np.array_equal(mask, g["ids_mask"])


def test_slots_pairs_and_targets(small):
    _, _, _, g, tgt = small
    assert np.array_equal(tgt["med_mask"], g["med_mask"])
    np.testing.assert_allclose(tgt["med_raw"][tgt["med_mask"]], g["med_raw_vals"], atol=2e-6)
    assert np.array_equal(tgt["pair"], g["pair"])
    assert np.array_equal(tgt["tgt_low_mask"], g["tgt_low_mask"])
    assert np.array_equal(tgt["tgt_med_mask"], g["tgt_med_mask"])
    np.testing.assert_allclose(tgt["tgt_low"][tgt["tgt_low_mask"]], g["tgt_low_vals"], atol=5e-5)
    np.testing.assert_allclose(tgt["tgt_med"][tgt["tgt_med_mask"]], g["tgt_med_vals"], atol=5e-5)
    np.testing.assert_allclose(tgt["tgt_top"], g["tgt_top"], atol=5e-5)


def test_normal_and_curvature(small):
    _, _, _, g, tgt = small
    m = g["ids_mask"]
    np.testing.assert_allclose(tgt["curvature"], g["curvature"], rtol=1e-4, atol=1e-7)
    s = tgt["singular"]
    well = (s[:, 1] - s[:, 2]) > 1e-4 * np.maximum(s[:, 0], 1e-12)
    mine = align_sign(g["normal"], tgt["normal"])
    assert np.abs(mine[well] - g["normal"][well]).max() < 1e-3
    np.testing.assert_allclose(align_sign(g["tgt_normal"], tgt["tgt_normal"])[well[m]],
                               g["tgt_normal"][well[m]], atol=1e-3)


def test_window_bookkeeping(small):
    _, cfg, _, g, tgt = small
    rows = tgt["pillar_coors"]
    coors = np.concatenate([rows[g["ids_keep"]], rows[g["ids_mask"]]])
    for s in (0, 1):
        win, ciw = O.window_partition(coors, cfg, s)
        lvl, _ = O.window_levels(win, cfg)
        assert np.array_equal(win, g[f"dec_win_shift{s}"])
        assert np.array_equal(ciw, g[f"dec_ciw_shift{s}"])
        assert np.array_equal(lvl, g[f"dec_lvl_shift{s}"])
        assert set(np.unique(lvl)) == {0, 1}, "fixture must exercise both buckets"


def _run(name, with_grad=True):
    case, cfg, frames, g = load_case(name)
    params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, case["param_seed"]).items()}
    trace = {}
    losses, pred, tgt = O.forward_train(params, frames, cfg, g["ids_keep"], g["ids_mask"], trace=trace)
    if with_grad:
        # sign-align the oracle's normal targets to the golden's before comparing the normal loss
        sum(losses.values()).backward()
    return case, cfg, g, params, losses, pred, tgt, trace


def _check_losses(g, losses, tol):
    for k, v in losses.items():
        ref = float(g["loss/" + k])
        assert abs(float(v) - ref) <= tol * abs(ref), (k, float(v), ref)


def test_small_forward_backward():
    case, cfg, g, params, losses, pred, tgt, trace = _run("small_b2")
    np.testing.assert_allclose(trace["voxel_features"].detach().numpy()[::8], g["voxel_features_rows8"],
                               rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pred["reg_top"].detach().numpy(), g["pred_reg_top"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(pred["nor_top"].detach().numpy(), g["pred_nor_top"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(pred["reg_med"].detach().numpy(), g["pred_reg_med"], rtol=1e-3, atol=2e-4)
    np.testing.assert_allclose(pred["cls_med"].detach().numpy(), g["pred_cls_med"], rtol=1e-3, atol=2e-4)
    _check_losses(g, losses, 1e-4)
    for k, p in params.items():
        ref = float(g["gradnorm/" + k])
        got = float(p.grad.double().norm())
        assert abs(got - ref) <= 2e-3 * max(ref, 1e-6) + 1e-7, (k, got, ref)
    for k in [k for k in g if k.startswith("grad/")]:
        np.testing.assert_allclose(params[k[5:]].grad.numpy(), g[k], rtol=5e-3, atol=1e-5)


def test_config0_single_frame_one_block():
    _, _, g, _, losses, *_ = _run("config0_1frame_1block", with_grad=False)
    _check_losses(g, losses, 1e-4)


def test_full_config_two_frames():
    _, _, g, params, losses, *_ = _run("full_b2")
    _check_losses(g, losses, 1e-4)
    bad = []
    for k, p in params.items():
        ref = float(g["gradnorm/" + k])
        got = float(p.grad.double().norm())
        if abs(got - ref) > 5e-3 * max(ref, 1e-6) + 1e-7:
            bad.append((k, got, ref))
    assert not bad, bad[:5]


@pytest.mark.parametrize("name", ["waymo_b2", "dense_b1"])
def test_other_geometries(name):
    """BASELINE.json configs[3] / configs[4] shapes: Waymo-shaped 0.32 m pillars on a 468 x 468 grid and the dense-grid
    stress (0.1 m pillars, 1024 x 1024).  The unmodified reference ran with only range / voxel sizes / grid replaced
    in its own config (oracle/make_golden.py::apply_geometry)."""
    case, cfg, g, params, losses, pred, tgt, trace = _run(name)
    assert trace["voxel_features"].shape[0] == int(g["n_pillars"])
    got = float(trace["voxel_features"].detach().double().abs().sum())
    assert abs(got - float(g["voxel_features_absum"])) <= 1e-4 * float(g["voxel_features_absum"])
    _check_losses(g, losses, 1e-4)
    bad = []
    for k, p in params.items():
        ref = float(g["gradnorm/" + k])
        got = float(p.grad.double().norm())
        if abs(got - ref) > 5e-3 * max(ref, 1e-6) + 1e-7:
            bad.append((k, got, ref))
    assert not bad, bad[:5]
