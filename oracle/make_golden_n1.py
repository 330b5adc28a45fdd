"""TEST INFRASTRUCTURE — golden vectors for the fine-tune consumer.

Runs the UNMODIFIED reference ``SSTInputLayer`` (shuffle_voxels=False) and ``SSTSecondPretrainedv1`` on CPU through
``oracle/ref_harness.py`` on seeded synthetic pillars, checks the numpy/torch restatement in ``oracle/geomae_oracle.py``
against them on the spot, and writes ``tests/golden/finetune_b2.npz``.  Needs /root/reference; the committed fixture
is what travels.

    python -m oracle.make_golden_n1
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geomae_b200.synthetic import make_frame  # noqa: E402
from oracle import geomae_oracle as O  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

# geometry of configs/pre_sst/m_sst_nus_second_pointpillar_fpn355_222_curv_07_ssl_data_wo_dbsampler_6x_1e-5.py:15-17,34
# with SMALL buckets so that both drop stages really drop voxels on a 2-frame batch
CASE = dict(
    frames=(dict(seed=11, point_scale=0.5), dict(seed=12, point_scale=0.25, sweeps=3)), feat_seed=5, param_seed=3,
    pc_range=(-50.0, -50.0, -5.0, 50.0, 50.0, 3.0), voxel_size=(0.25, 0.25, 8), window_shape=(12, 12),
    shifts=((0, 0), (6, 6)),
    drop_info={0: dict(max_tokens=8, drop_range=(0, 8)), 1: dict(max_tokens=20, drop_range=(8, 20)),
               2: dict(max_tokens=36, drop_range=(20, 100000))},
    n_blocks=1, output_shape=(400, 400), conv_in=128, conv_out=(16, 16, 32), layer_nums=(1, 1, 1), strides=(2, 2, 2))


def case_cfg(case=CASE):
    return O.PathConfig(pc_range=case["pc_range"], voxel_size=case["voxel_size"], window_shape=case["window_shape"],
                        shifts=case["shifts"], drop_info=case["drop_info"], grid_size=(1, 400, 400))


def case_inputs(case=CASE):
    cfg = case_cfg(case)
    frames = [make_frame(**kw) for kw in case["frames"]]
    coors, _, _ = O.unique_rows(O.batch_voxelize(frames, cfg.voxel_size, cfg.pc_range))
    g = torch.Generator().manual_seed(case["feat_seed"])
    feat = torch.randn(coors.shape[0], cfg.d_model, generator=g)
    return cfg, coors, feat


def build_reference(case=CASE):
    cfg = case_cfg(case)
    InputLayer, Second = H.load_finetune_consumer()
    layer = InputLayer(drop_info=case["drop_info"], shifts_list=list(case["shifts"]), window_shape=case["window_shape"],
                       point_cloud_range=list(case["pc_range"]), voxel_size=tuple(case["voxel_size"]),
                       shuffle_voxels=False, debug=True)
    nb = max(2, case["n_blocks"])        # the reference indexes d_model[1] (:285)
    bb = Second(d_model=[cfg.d_model] * nb, nhead=[cfg.nhead] * nb, num_blocks=case["n_blocks"],
                dim_feedforward=[cfg.ffn] * nb, output_shape=list(case["output_shape"]),
                conv_in_channels=case["conv_in"], conv_out_channels=list(case["conv_out"]),
                layer_nums=list(case["layer_nums"]), layer_strides=list(case["strides"]), drop_info=case["drop_info"],
                window_shape=case["window_shape"], debug=True)
    params = O.init_params_second(cfg, case["n_blocks"], case["conv_in"], case["conv_out"], case["layer_nums"],
                                  case["param_seed"])
    sd = bb.state_dict()
    stripped = {k[len("backbone."):]: v for k, v in params.items()}
    assert not [k for k in stripped if k not in sd], [k for k in stripped if k not in sd]
    bb.load_state_dict({**sd, **stripped})
    layer.train(), bb.train()
    return layer, bb, params


def main():
    case = CASE
    cfg, coors, feat = case_inputs(case)
    layer, bb, params = build_reference(case)
    n_frames = len(case["frames"])

    # (A) the drop itself.  The reference ranks voxels inside a window with an UNSTABLE torch.sort, so which voxels
    # survive is implementation-defined; pin the restatement by giving it the reference's own rank function — the
    # bucket rule, the budgets and the two-stage survivor logic must then agree index for index.
    _, _, info_a = layer(feat, torch.from_numpy(coors), n_frames)
    keep_a, levels_a = O.input_layer_drop(
        coors, cfg, inner_fn=lambda w: layer.get_inner_win_inds(torch.from_numpy(w)).numpy())
    assert np.array_equal(keep_a, info_a["voxel_keep_inds"].numpy()), "keep indices differ"
    for i in range(2):
        assert np.array_equal(levels_a[i], info_a[f"voxel_drop_level_shift{i}"].numpy()), f"levels of shift {i} differ"
    print(f"(A) {coors.shape[0]} pillars, reference keeps {keep_a.size}: restatement with the reference's rank "
          f"function agrees index for index")

    # (B) everything downstream, on the survivors of the STABLE-rank drop (what the CUDA path computes with seed 0).
    # A second pass over survivors drops nothing (every window already fits its bucket), so the reference's output
    # no longer depends on its sort order and can be compared exactly.
    keep, _ = O.input_layer_drop(coors, cfg)
    coors_k = np.ascontiguousarray(coors[keep])
    feat_k = feat[torch.from_numpy(keep)].clone().requires_grad_(True)
    tup = layer(feat_k, torch.from_numpy(coors_k), n_frames)
    kept_feat, inds_list, info = tup
    assert info["voxel_keep_inds"].numel() == keep.size, "second pass dropped voxels"
    outs = bb(tup)
    loss = sum((o * o).mean() for o in outs)
    loss.backward()
    grads = {"backbone." + k: p.grad for k, p in bb.named_parameters()}

    g = dict(coors=coors.astype(np.int16), n_pillars=np.int64(coors.shape[0]),
             feat_absum=np.float64(feat.double().abs().sum().item()), keep_inds=keep.astype(np.int32),
             ref_unstable_keep_count=np.int64(keep_a.size), loss=np.float64(loss.item()),
             d_feat_norm=np.float64(feat_k.grad.double().norm().item()), d_feat_rows8=feat_k.grad.numpy()[::8])
    for i in range(2):
        g[f"level_shift{i}"] = info[f"voxel_drop_level_shift{i}"].numpy().astype(np.int8)
        g[f"batch_win_inds_shift{i}"] = info[f"batch_win_inds_shift{i}"].numpy().astype(np.int32)
        g[f"coors_in_win_shift{i}"] = info[f"coors_in_win_shift{i}"].numpy().astype(np.int8)
        for dl, (slot, where) in inds_list[i].items():
            mt = case["drop_info"][dl]["max_tokens"]
            g[f"win_slot_shift{i}_level{dl}"] = (slot // mt).numpy().astype(np.int32)
            g[f"where_shift{i}_level{dl}"] = where[0].numpy().astype(np.int32)
    for i, o in enumerate(outs):
        g[f"out{i}_shape"] = np.array(o.shape, np.int64)
        g[f"out{i}_absum"] = np.float64(o.detach().double().abs().sum().item())
        g[f"out{i}_sub"] = o.detach().numpy()[:, :, ::5, ::5]
    for k, v in grads.items():
        g["gradnorm/" + k] = np.float64(v.double().norm().item())

    # the restatement must reproduce (B) before the fixture is written
    keep2, levels, layout, enc, o_outs = O.sst_second_forward(
        params, feat_k.detach(), coors_k, n_frames, cfg, case["n_blocks"], case["output_shape"], case["layer_nums"],
        case["strides"])
    assert keep2.size == keep.size
    for i in range(2):
        assert np.array_equal(levels[i], g[f"level_shift{i}"]), f"levels of shift {i} differ"
        assert np.array_equal(layout.shifts[i]["win"], g[f"batch_win_inds_shift{i}"])
        assert np.array_equal(layout.shifts[i]["ciw"], g[f"coors_in_win_shift{i}"])
    for i, (a, b) in enumerate(zip(o_outs, outs)):
        err = (a - b.detach()).abs().max().item()
        assert err < 1e-4, (i, err)
        print(f"(B) stage {i}: restatement vs reference max abs err {err:.2e}, shape {tuple(b.shape)}")
    print(f"(B) {coors.shape[0]} pillars, {coors.shape[0] - keep.size} dropped; levels shift0 "
          f"{np.bincount(levels[0])}, shift1 {np.bincount(levels[1])}")
    out = os.path.join(ROOT, "tests", "golden", "finetune_b2.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
