"""GPU parity of the fine-tune consumer against the oracle and the reference-made golden:
SSTInputLayer (window drop kernel), recover_bev, SSTSecondPretrainedv1, DynamicVoxelNet + checkpoint loading by key."""
import numpy as np
import pytest
import torch

from oracle import geomae_oracle as O
from oracle.make_golden_n1 import CASE
from tests.test_finetune_cpu import budget_of, check_drop_properties, load

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def input_layer(shuffle, drop_info=None):
    from geomae_b200.sst_input_layer import SSTInputLayer
    return SSTInputLayer(drop_info=drop_info or CASE["drop_info"], shifts_list=list(CASE["shifts"]),
                         window_shape=CASE["window_shape"], point_cloud_range=list(CASE["pc_range"]),
                         voxel_size=CASE["voxel_size"], shuffle_voxels=shuffle, debug=True).to(DEV)


def check_flat2win(inds, levels, win, drop_info):
    """window slot = rank of the window among its level's windows; in-window slots form 0..n_w-1."""
    for dl, (slot, where) in inds.items():
        mt = drop_info[dl]["max_tokens"]
        slot, sel = slot.cpu().numpy(), where[0].cpu().numpy()
        assert np.array_equal(sel, np.where(levels == dl)[0])
        uniq, rank = np.unique(win[sel], return_inverse=True)
        assert np.array_equal(slot // mt, rank)
        inner = slot % mt
        order = np.lexsort((inner, rank))
        starts = np.searchsorted(rank[order], np.arange(uniq.size))
        assert np.array_equal(inner[order], np.arange(sel.size) - starts[rank[order]])


def test_input_layer_stable_drop_is_bit_exact():
    g, cfg, coors, feat = load()
    layer = input_layer(shuffle=False)
    out_feat, inds_list, info = layer(feat.to(DEV), torch.from_numpy(coors).to(DEV), len(CASE["frames"]))
    keep, levels = O.input_layer_drop(coors, cfg)
    assert np.array_equal(info["voxel_keep_inds"].cpu().numpy(), keep)
    assert np.array_equal(info["voxel_keep_inds"].cpu().numpy(), g["keep_inds"])
    assert info["coors"].dtype == torch.int64 and np.array_equal(info["coors"].cpu().numpy(), coors[keep])
    assert torch.equal(out_feat.cpu(), feat[torch.from_numpy(keep)])
    for i in range(2):
        lv = info[f"voxel_drop_level_shift{i}"].cpu().numpy()
        assert np.array_equal(lv, levels[i])
        win, ciw = O.window_partition(coors[keep], cfg, i)
        assert np.array_equal(info[f"batch_win_inds_shift{i}"].cpu().numpy(), win)
        assert np.array_equal(info[f"coors_in_win_shift{i}"].cpu().numpy(), ciw)
        check_flat2win(inds_list[i], lv, win, cfg.drop_info)
    # a second pass over survivors is the golden's case (B): nothing dropped, the reference's own levels
    _, inds2, info2 = layer(out_feat, info["coors"], len(CASE["frames"]))
    assert info2["voxel_keep_inds"].numel() == keep.size
    for i in range(2):
        assert np.array_equal(info2[f"voxel_drop_level_shift{i}"].cpu().numpy(), g[f"level_shift{i}"])
        for dl, (slot, where) in inds2[i].items():
            assert np.array_equal((slot // cfg.drop_info[dl]["max_tokens"]).cpu().numpy(), g[f"win_slot_shift{i}_level{dl}"])
            assert np.array_equal(where[0].cpu().numpy(), g[f"where_shift{i}_level{dl}"])


def test_input_layer_shuffled_drop_is_a_valid_random_instance():
    _, cfg, coors, feat = load()
    layer = input_layer(shuffle=True)
    kept = []
    for seed in (1, 2):
        torch.manual_seed(seed)
        _, _, info = layer(feat.to(DEV), torch.from_numpy(coors).to(DEV), len(CASE["frames"]))
        keep = info["voxel_keep_inds"].cpu().numpy()
        levels = [info[f"voxel_drop_level_shift{i}"].cpu().numpy() for i in range(2)]
        check_drop_properties(coors, keep, levels, cfg)
        kept.append(keep)
    stable, _ = O.input_layer_drop(coors, cfg)
    assert abs(kept[0].size - stable.size) < 0.02 * stable.size
    assert not np.array_equal(kept[0], kept[1]) and not np.array_equal(kept[0], stable[:kept[0].size])
    # the subset is not biased towards low token indices: survivors of over-full shift-0 windows are spread evenly
    w0 = O.window_partition(coors, cfg, 0)[0]
    cnt = np.bincount(w0)
    _, budget = budget_of(cnt, cfg.drop_info)
    over = (cnt > budget)[w0]
    rank_in_win = O.inner_win_inds_stable(w0)
    rel = (rank_in_win[over] / cnt[w0][over])
    kept_mask = np.zeros(coors.shape[0], bool)
    kept_mask[kept[0]] = True
    assert abs(rel[kept_mask[over]].mean() - rel.mean()) < 0.03


def test_input_layer_fine_tune_config_never_drops():
    """configs/pre_sst/…6x_1e-5.py:16-30: 12x12 windows with a 144-token top bucket — the drop is the identity."""
    _, cfg, coors, feat = load()
    di = {0: dict(max_tokens=32, drop_range=(0, 32)), 1: dict(max_tokens=72, drop_range=(32, 72)),
          2: dict(max_tokens=144, drop_range=(72, 1000))}
    layer = input_layer(shuffle=True, drop_info=(di, di))
    out_feat, _, info = layer(feat.to(DEV), torch.from_numpy(coors).to(DEV), len(CASE["frames"]))
    assert info["voxel_keep_inds"].numel() == coors.shape[0] and out_feat.shape[0] == coors.shape[0]
    cnt = np.bincount(O.window_partition(coors, cfg, 0)[0])
    lvl, _ = budget_of(cnt, di)
    assert np.array_equal(info["voxel_drop_level_shift0"].cpu().numpy(), lvl[O.window_partition(coors, cfg, 0)[0]])


def test_recover_bev_forward_backward():
    from geomae_b200.sst_second import _RecoverBEV
    _, cfg, coors, _ = load()
    torch.manual_seed(0)
    for c in (128, 24):
        feat = torch.randn(coors.shape[0], c, device=DEV, requires_grad=True)
        canvas = _RecoverBEV.apply(feat, torch.from_numpy(coors).to(DEV).int(), 2, 400, 400)
        ref = O.recover_bev(feat.detach().cpu(), coors, 2, 400, 400)
        assert torch.equal(canvas.cpu(), ref)
        w = torch.randn_like(canvas)
        (canvas * w).sum().backward()
        b, y, x = (torch.from_numpy(coors[:, i].astype(np.int64)).to(DEV) for i in (0, 2, 3))
        assert torch.equal(feat.grad, w[b, :, y, x])


def build_second():
    from geomae_b200.sst_second import SSTSecondPretrainedv1
    nb = CASE["n_blocks"]
    return SSTSecondPretrainedv1(d_model=[128] * nb, nhead=[8] * nb, num_blocks=nb, dim_feedforward=[256] * nb,
                                 output_shape=list(CASE["output_shape"]), conv_in_channels=CASE["conv_in"],
                                 conv_out_channels=list(CASE["conv_out"]), layer_nums=list(CASE["layer_nums"]),
                                 layer_strides=list(CASE["strides"]), drop_info=CASE["drop_info"],
                                 window_shape=CASE["window_shape"], debug=True).to(DEV)


@pytest.mark.parametrize("impl,out_tol,loss_tol,grad_tol", (("tc3", 2e-4, 2e-4, 5e-3), ("tc1", 6e-2, 1e-2, 5e-2)))
def test_second_backbone_matches_reference_golden(impl, out_tol, loss_tol, grad_tol):
    """forward + backward of SSTInputLayer -> SSTSecondPretrainedv1 against the unmodified reference's outputs.
    tc3 (bf16x3 split, fp32-equivalent): 2e-4 abs on the BN-normalised stage outputs (|x| up to ~5);
    tc1 (plain bf16 operands, 2^-9 relative per GEMM, amplified by the batch norms): 6e-2 abs, 1 % on the loss."""
    g, cfg, coors, feat = load()
    bb = build_second()
    bb.set_sra_impl(impl)
    params = O.init_params_second(cfg, CASE["n_blocks"], CASE["conv_in"], CASE["conv_out"], CASE["layer_nums"],
                                  CASE["param_seed"])
    sd = bb.state_dict()
    bb.load_state_dict({**sd, **{k[len("backbone."):]: v for k, v in params.items()}})
    bb.train()
    layer = input_layer(shuffle=False)
    x = feat.to(DEV).requires_grad_(True)
    tup = layer(x, torch.from_numpy(coors).to(DEV), len(CASE["frames"]))
    outs = bb(tup)
    loss = sum((o * o).mean() for o in outs)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = impl == "tc1"       # convolution backward in fp32 for the parity mode
    try:
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert abs(loss.item() - g["loss"]) < loss_tol * abs(g["loss"])
    for i, o in enumerate(outs):
        assert tuple(o.shape) == tuple(g[f"out{i}_shape"])
        err = np.abs(o.detach().cpu().numpy()[:, :, ::5, ::5] - g[f"out{i}_sub"]).max()
        assert err < out_tol, (i, err)
    keep = torch.from_numpy(g["keep_inds"].astype(np.int64))
    d_feat = x.grad.cpu()
    mask = torch.ones(x.shape[0], dtype=torch.bool)
    mask[keep] = False
    assert (d_feat[mask] == 0).all()                       # dropped voxels receive no gradient
    ref_rows = g["d_feat_rows8"]
    got = d_feat[keep].numpy()[::8]
    assert np.abs(got - ref_rows).max() < 10 * grad_tol * np.abs(ref_rows).max()
    for k, p in bb.named_parameters():
        ref = g["gradnorm/backbone." + k]
        assert abs(p.grad.double().norm().item() - ref) <= grad_tol * ref + 1e-9, k


def test_dynamic_voxelnet_loads_pretraining_checkpoint_by_key(tmp_path):
    """A checkpoint written by the pre-training detector loads into the fine-tune consumer by key (what
    `load_from = …/epoch_72.pth` does, configs/pre_sst/…6x_1e-5.py:280): VFE + encoder blocks are taken over, and the
    consumer's encoder then computes exactly what the pre-training backbone's encoder computes."""
    import os
    import geomae_b200 as G
    from geomae_b200.registry import Config
    from geomae_b200.synthetic import make_frame
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mae = Config.fromfile(os.path.join(root, "configs/mae_sst/geomae_nus_pretrain.py"))
    torch.manual_seed(0)
    pre = G.build_detector(mae.model).to(DEV)
    ckpt = tmp_path / "epoch_72.pth"
    torch.save(dict(state_dict={("module." + k): v for k, v in pre.state_dict().items()}, meta=dict(epoch=72)), ckpt)

    vs, rng, win = (0.256, 0.256, 8), [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], (12, 12)
    di = {0: dict(max_tokens=32, drop_range=(0, 32)), 1: dict(max_tokens=72, drop_range=(32, 72)),
          2: dict(max_tokens=144, drop_range=(72, 1000))}
    model = dict(
        type="DynamicVoxelNet",
        voxel_layer=dict(voxel_size=vs, max_num_points=-1, point_cloud_range=rng, max_voxels=(-1, -1)),
        voxel_encoder=dict(type="DynamicScatterVFE", in_channels=5, feat_channels=[64, 128], with_distance=False,
                           voxel_size=vs, with_cluster_center=True, with_voxel_center=True, point_cloud_range=rng,
                           norm_cfg=dict(type="naiveSyncBN1d", eps=1e-3, momentum=0.01)),
        middle_encoder=dict(type="SSTInputLayer", window_shape=win, shifts_list=[(0, 0), (6, 6)], point_cloud_range=rng,
                            voxel_size=vs, shuffle_voxels=True, debug=True, drop_info=(di, di)),
        backbone=dict(type="SSTSecondPretrainedv1", d_model=[128] * 6, nhead=[8] * 6, num_blocks=6,
                      dim_feedforward=[256] * 6, output_shape=[400, 400], conv_in_channels=128,
                      conv_out_channels=[32, 32, 64], layer_nums=[1, 1, 1], layer_strides=[2, 2, 2], debug=True,
                      drop_info=(di, di), pos_temperature=10000, normalize_pos=False, window_shape=win),
        neck=dict(type="SECONDFPN"), bbox_head=dict(type="Anchor3DHead"))
    det = G.build_detector(model).to(DEV)
    loaded, untouched, unexpected = det.load_pretrained(str(ckpt))
    assert any(k.startswith("voxel_encoder.") for k in loaded)
    enc_keys = [k for k in det.state_dict() if k.startswith("backbone.encoder_blocks.")]
    assert enc_keys and set(enc_keys) <= set(loaded)
    assert all(k.startswith("backbone.conv_blocks.") for k in untouched)
    assert all(k.split(".")[1].startswith(("decoder", "mask_token", "cls_pred")) for k in unexpected if k.startswith("backbone."))

    frames = [torch.from_numpy(make_frame(seed=s)).to(DEV) for s in (3, 4)]
    det.train(), pre.train()
    outs = det.extract_feat(frames)
    assert [tuple(o.shape) for o in outs] == [(2, 32, 200, 200), (2, 32, 100, 100), (2, 64, 50, 50)]
    assert all(torch.isfinite(o).all() for o in outs)
    # same pillars, same weights: the consumer's encoder output == the pre-training backbone's encoder output
    from geomae_b200.voxel import scatter_frames
    from geomae_b200.windows import WindowLayout
    pb = scatter_frames(pre.geom, frames)
    feats, fcoors = pre.voxel_encoder(pb)
    enc_pre = pre.backbone.forward_encoder(feats, WindowLayout.from_coors(pre.backbone.spec, pre.backbone.geom, fcoors, 2))
    tup = det.middle_encoder(feats, fcoors, 2)
    assert tup[2]["voxel_keep_inds"].numel() == fcoors.shape[0]
    bb = det.backbone
    from geomae_b200.windows import pos_table
    enc_ft = bb._stack(2)(tup[0], tup[2]["window_layout"], pos_table(win, 128, 10000, DEV), 3)
    assert torch.equal(enc_pre, enc_ft)
    with pytest.raises(NotImplementedError):
        det.forward_train(points=frames)
