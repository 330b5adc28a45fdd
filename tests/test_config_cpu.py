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
