"""Oracle of the fine-tune consumer (SSTInputLayer voxel drop + SSTSecondPretrainedv1) against the
golden vectors the unmodified reference produced (oracle/make_golden_n1.py)."""
import os

import numpy as np
import torch

from oracle import geomae_oracle as O
from oracle.make_golden_n1 import CASE, case_cfg, case_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "finetune_b2.npz")


def load():
    g = dict(np.load(GOLDEN))
    cfg, coors, feat = case_inputs(CASE)
    assert np.array_equal(coors, g["coors"]), "synthetic generator drifted from the committed golden inputs"
    assert abs(feat.double().abs().sum().item() - g["feat_absum"]) < 1e-6 * g["feat_absum"]
    return g, cfg, coors, feat


def budget_of(n, drop_info):
    """bucket rule lower < n <= upper (middle_encoders/sst_input_layer.py:222) -> (level, max_tokens)."""
    lvl, budget = np.full(n.shape, -1), np.zeros(n.shape, np.int64)
    for dl, info in drop_info.items():
        lo, hi = info["drop_range"]
        m = (n > lo) & (n <= hi)
        lvl[m], budget[m] = dl, info["max_tokens"]
    return lvl, budget


def check_drop_properties(coors, keep, levels, cfg):
    """What every valid instance of the reference's drop satisfies, whatever rank order its sort produced."""
    n = coors.shape[0]
    w0 = O.window_partition(coors, cfg, 0)[0]
    cnt0 = np.bincount(w0)
    lvl0_w, budget0_w = budget_of(cnt0, cfg.drop_info)
    kept0 = np.bincount(w0[keep], minlength=cnt0.size)
    assert (kept0 <= budget0_w).all()                                   # never above the bucket's budget
    assert np.array_equal(levels[0], lvl0_w[w0[keep]])                  # level = bucket of the PRE-drop count
    w1 = O.window_partition(coors, cfg, 1)[0]
    kept1 = np.bincount(w1[keep], minlength=w1.max() + 1)
    # a shift-1 window's level comes from its count among stage-0 survivors (>= what it finally keeps, <= its full count)
    lvl1_lo, _ = budget_of(kept1, cfg.drop_info)
    lvl1_hi, _ = budget_of(np.bincount(w1, minlength=kept1.size), cfg.drop_info)
    l1 = levels[1]
    assert (l1 >= lvl1_lo[w1[keep]]).all() and (l1 <= lvl1_hi[w1[keep]]).all()
    budget1 = np.array([cfg.drop_info[int(l)]["max_tokens"] for l in l1])
    assert (kept1[w1[keep]] <= budget1).all()
    # nothing is dropped without cause: total kept >= sum over shift-0 windows of min(n, budget) minus stage-1 cuts
    assert keep.size <= np.minimum(cnt0, budget0_w).sum()
    assert np.array_equal(np.sort(keep), keep) and keep.size == np.unique(keep).size and keep.max() < n


def test_drop_restatement_matches_golden_and_properties():
    g, cfg, coors, _ = load()
    keep, levels = O.input_layer_drop(coors, cfg)
    assert np.array_equal(keep, g["keep_inds"])
    assert keep.size < coors.shape[0], "the case must actually drop voxels"
    check_drop_properties(coors, keep, levels, cfg)
    # the reference's own (unstable-sort) instance kept a different subset of almost the same size
    assert abs(int(g["ref_unstable_keep_count"]) - keep.size) < 0.01 * keep.size
    # a second pass over the survivors drops nothing and yields the golden levels
    coors_k = coors[keep]
    keep2, levels2 = O.input_layer_drop(coors_k, cfg)
    assert keep2.size == keep.size
    for i in range(2):
        assert np.array_equal(levels2[i], g[f"level_shift{i}"])
        win, ciw = O.window_partition(coors_k, cfg, i)
        assert np.array_equal(win, g[f"batch_win_inds_shift{i}"])
        assert np.array_equal(ciw, g[f"coors_in_win_shift{i}"])
        inds = O.flat2win_indices(win, levels2[i], cfg)
        for dl, (slot, sel, _) in inds.items():
            assert np.array_equal(sel, g[f"where_shift{i}_level{dl}"])
            assert np.array_equal(slot // cfg.drop_info[dl]["max_tokens"], g[f"win_slot_shift{i}_level{dl}"])


def test_bucket_rule_edges():
    """lower < n <= upper: a window with exactly `upper` voxels stays in the lower bucket (:222)."""
    cfg = case_cfg()
    for n, want in ((1, 0), (8, 0), (9, 1), (20, 1), (21, 2), (36, 2), (60, 2)):
        # n voxels in one 12x12 window of frame 0
        ys, xs = np.divmod(np.arange(n), 12)
        coors = np.stack([np.zeros(n, np.int64), np.zeros(n, np.int64), ys, xs], axis=1)
        keep, levels = O.input_layer_drop(coors, cfg)
        assert (levels[0] == want).all()
        assert keep.size == min(n, cfg.drop_info[want]["max_tokens"])
        assert np.array_equal(keep, np.arange(keep.size))     # stable rank: the first voxels stay


def test_second_restatement_matches_golden():
    g, cfg, coors, feat = load()
    params = O.init_params_second(cfg, CASE["n_blocks"], CASE["conv_in"], CASE["conv_out"], CASE["layer_nums"],
                                  CASE["param_seed"])
    keep = g["keep_inds"].astype(np.int64)
    x = feat[torch.from_numpy(keep)].clone().requires_grad_(True)
    for p in params.values():
        p.requires_grad_(True)
    _, _, _, _, outs = O.sst_second_forward(params, x, coors[keep], len(CASE["frames"]), cfg, CASE["n_blocks"],
                                            CASE["output_shape"], CASE["layer_nums"], CASE["strides"])
    loss = sum((o * o).mean() for o in outs)
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-5 * abs(g["loss"])
    for i, o in enumerate(outs):
        assert tuple(o.shape) == tuple(g[f"out{i}_shape"])
        np.testing.assert_allclose(o.detach().numpy()[:, :, ::5, ::5], g[f"out{i}_sub"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(x.grad.numpy()[::8], g["d_feat_rows8"], atol=1e-7, rtol=2e-3)
    for k, p in params.items():
        ref = g["gradnorm/" + k]
        assert abs(p.grad.double().norm().item() - ref) <= 2e-3 * ref + 1e-9, k
