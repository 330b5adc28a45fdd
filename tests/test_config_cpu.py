"""Config / registry surface: the reference's own config file builds the model unchanged."""
import os

import pytest
import torch

import geomae_b200  # noqa: F401
from geomae_b200.registry import Config, build_model
from oracle import geomae_oracle as O

REF_CFG = "/root/reference/configs/mae_sst/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5.py"
REF_CFG_8X = "/root/reference/configs/mae_sst/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_8x_1e-5.py"
OWN_CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs/mae_sst/geomae_nus_pretrain.py")


def test_own_config_builds_with_reference_state_dict_keys():
    model = build_model(Config.fromfile(OWN_CFG).model)
    assert sum(p.numel() for p in model.parameters()) == 2760854
    sd = model.state_dict()
    ref = O.init_params(O.PathConfig(), 0)
    assert not [k for k in ref if k not in sd]
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    # every parameter of the model is named in the oracle's (= the reference's) parameter tree
    assert not [k for k, _ in model.named_parameters() if k not in ref]
    # the reference xavier-initialises the [1,128] mask token too (…top_only.py:114,131,318-321)
    assert torch.count_nonzero(sd["backbone.mask_token"]) > 0


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
@pytest.mark.parametrize("path", [REF_CFG, REF_CFG_8X])
def test_reference_config_loads_unchanged(path):
    cfg = Config.fromfile(path)
    model = build_model(cfg.model)
    assert type(model).__name__ == "MultiSubVoxelDynamicVoxelNetSSL"
    assert type(model.backbone).__name__ == "MultiMAESSTSPChoose"
    assert type(model.voxel_encoder.vfe_layers[0].norm).__name__ == "NaiveSyncBatchNorm1d"
    assert cfg.data["samples_per_gpu"] == 4


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
def test_own_config_equals_reference_model_dict():
    a, b = Config.fromfile(OWN_CFG).model, Config.fromfile(REF_CFG).model

    def norm(x):
        if isinstance(x, dict):
            return {k: norm(v) for k, v in x.items()}
        if isinstance(x, (list, tuple)):
            return [norm(v) for v in x]
        return x
    assert norm(a) == norm(b)


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present")
def test_checkpoint_compatibility_with_the_unmodified_reference_detector():
    """A checkpoint written by the reference (`epoch_*.pth` = its state_dict) loads strictly into this model and the
    other way round: same keys (parameters AND buffers), same shapes, same dtypes."""
    from oracle import ref_harness as H
    det = H.build_detector(seed=0)
    ref_sd = det.state_dict()
    model = build_model(Config.fromfile(OWN_CFG).model)
    own_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(own_sd.keys())
    for k, v in ref_sd.items():
        assert tuple(v.shape) == tuple(own_sd[k].shape) and v.dtype == own_sd[k].dtype, k
    model.load_state_dict(ref_sd, strict=True)
    det.load_state_dict(model.state_dict(), strict=True)
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k


def test_cyclic_lr_matches_the_schedule_of_cosine_2x():
    """configs/_base_/schedules/cosine_2x.py:10-15: 1 -> 100x over the first 10 % of the run, then down to 1e-3x (cosine)."""
    from geomae_b200.train import cyclic_lr
    base, total = 1e-5, 1000
    assert abs(cyclic_lr(base, 0, total) - base) < 1e-12
    assert abs(cyclic_lr(base, 100, total) - 100 * base) < 1e-9
    assert abs(cyclic_lr(base, 50, total) - base * (100 + 0.5 * (1 - 100) * 1.0)) < 1e-9      # half way up: cos(pi/2) + 1 = 1
    assert abs(cyclic_lr(base, 999, total) - base * 1e-3) < base * 1e-3
    assert all(cyclic_lr(base, i, total) >= cyclic_lr(base, i + 1, total) for i in range(100, 998))


def test_flat_trainer_layout_buckets():
    """Flat layout [decay early | decay late | no-decay late | no-decay early]: decayed prefix, contiguous late bucket,
    every parameter inside exactly one bucket (the early bucket is all-reduced from inside backward)."""
    import torch
    from geomae_b200.train import FlatTrainer

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = torch.nn.Module()
            self.backbone.decoder_pred_top = torch.nn.Linear(8, 3)
            self.backbone.decoder_centroid_blocks = torch.nn.ModuleList([torch.nn.Linear(8, 8), torch.nn.LayerNorm(8)])
            self.backbone.decoder_centroid_blocks[1].register_parameter("norm_w", torch.nn.Parameter(torch.ones(8)))
            self.backbone.encoder_blocks = torch.nn.ModuleList([torch.nn.Linear(8, 8)])
            self.backbone.norm1 = torch.nn.LayerNorm(8)
            self.voxel_encoder = torch.nn.Linear(5, 8)
    net = Net()
    tr = FlatTrainer(net)
    (a0, a1), (b0, b1) = tr.early_ranges
    l0, l1 = tr.late_range
    assert a0 == 0 and a1 == l0 and l1 == b0 and b1 == tr.n and l0 <= tr.n_decay <= l1
    off = 0
    for k, p in tr.order:
        early = k.startswith(FlatTrainer.EARLY_KEYS)
        inside_early = (a0 <= off < a1) or (b0 <= off < b1)
        assert early == inside_early, k
        assert ("norm" in k) == (off >= tr.n_decay), k
        assert p.data_ptr() == tr.flat_param.data_ptr() + 4 * off
        off += (p.numel() + 63) // 64 * 64
    assert off == tr.n


REF_FT_CFG = ("/root/reference/configs/pre_sst/"
              "m_sst_nus_second_pointpillar_fpn355_222_curv_07_ssl_data_wo_dbsampler_6x_1e-5.py")
OWN_FT_CFG = os.path.join(os.path.dirname(OWN_CFG), "..", "pre_sst", "geomae_nus_finetune_features.py")


def check_finetune_consumer(model, pre):
    assert type(model).__name__ == "DynamicVoxelNet"
    assert type(model.middle_encoder).__name__ == "SSTInputLayer" and model.middle_encoder.shuffle_voxels
    assert type(model.backbone).__name__ == "SSTSecondPretrainedv1"
    assert type(model.backbone.conv_blocks[0][1]).__name__ == "NaiveSyncBatchNorm2d"
    assert [len(b) for b in model.backbone.conv_blocks] == [12, 18, 18]       # (1 + layer_num) x (conv, BN, ReLU)
    # every VFE and encoder tensor of the pre-training checkpoint has a destination of the same shape (load by key)
    own, sd = model.state_dict(), pre.state_dict()
    shared = [k for k in sd if k in own]
    assert len([k for k in shared if k.startswith("backbone.encoder_blocks.")]) == 6 * 2 * 12
    assert len([k for k in shared if k.startswith("voxel_encoder.")]) == len([k for k in sd if k.startswith("voxel_encoder.")])
    assert all(own[k].shape == sd[k].shape for k in shared)
    loaded, untouched, unexpected = model.load_pretrained(dict(state_dict=sd))
    assert sorted(loaded) == sorted(shared)
    assert all(k.startswith("backbone.conv_blocks.") for k in untouched)
    for k in loaded:
        assert torch.equal(model.state_dict()[k], sd[k]), k


def test_own_finetune_config_builds_and_takes_the_pretraining_checkpoint():
    torch.manual_seed(0)
    pre = build_model(Config.fromfile(OWN_CFG).model)
    check_finetune_consumer(build_model(Config.fromfile(OWN_FT_CFG).model), pre)


@pytest.mark.skipif(not os.path.exists(REF_FT_CFG), reason="reference tree not present")
def test_reference_finetune_config_loads_unchanged():
    cfg = Config.fromfile(REF_FT_CFG)
    torch.manual_seed(0)
    pre = build_model(Config.fromfile(OWN_CFG).model)
    model = build_model(cfg.model, train_cfg=cfg.get("train_cfg"), test_cfg=cfg.get("test_cfg"))
    check_finetune_consumer(model, pre)
    assert model.bbox_head_cfg["type"] == "Anchor3DHead" and model.neck_cfg["type"] == "SECONDFPN"
    own = Config.fromfile(OWN_FT_CFG).model
    for part in ("voxel_layer", "voxel_encoder", "middle_encoder", "backbone"):
        a, b = dict(own[part]), dict(cfg.model[part])
        assert {k: (list(v) if isinstance(v, tuple) else v) for k, v in a.items()} == \
               {k: (list(v) if isinstance(v, tuple) else v) for k, v in b.items()}, part
