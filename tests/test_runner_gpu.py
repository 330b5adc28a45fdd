"""Files -> BatchLoader -> train_step_from_host (device augmentation) -> checkpoints -> resume, on synthetic nuScenes
files (SURVEY.md §8(f) N3 + §5 checkpoint/resume)."""
import os

import numpy as np
import pytest
import torch

from tests.test_dataset_cpu import TRAIN_PIPELINE, write_dataset

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build():
    import geomae_b200 as G
    from geomae_b200.registry import Config
    cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
    torch.manual_seed(0)
    model = G.build_detector(cfg.model).to(DEV)
    model.set_impl("tc3")
    return model


def test_epochs_checkpoint_and_resume(tmp_path):
    from geomae_b200.dataset import BatchLoader, NuScenesDatasetSSL
    from geomae_b200.runner import load_checkpoint, train
    ann = write_dataset(tmp_path, n_scenes=6)
    ds = NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE)

    def run(work, epochs, resume=None, seed=11, until=None, dataset=None):
        torch.manual_seed(seed)                 # mask-split seeds come from the CPU generator
        model = build()
        loader = BatchLoader(dataset or ds, samples_per_gpu=2, seed=1, workers=2)
        lines = []
        trainer, hist = train(model, loader, str(work), epochs, base_lr=1e-4, resume_from=resume, log_interval=2,
                              log=lines.append, until_epoch=until)
        return model, trainer, hist, lines

    m2, t2, h2, lines = run(tmp_path / "a", 2)
    assert len(h2) == 6 and all(np.isfinite(h2)) and len(lines) == 4 and "Epoch [2][3/3]" in lines[-1]
    assert sorted(os.listdir(tmp_path / "a")) == ["epoch_1.pth", "epoch_2.pth"]
    ckpt = torch.load(tmp_path / "a" / "epoch_1.pth", map_location="cpu", weights_only=False)
    assert set(ckpt) == {"meta", "state_dict", "optimizer"} and ckpt["meta"]["epoch"] == 1 and ckpt["meta"]["iter"] == 3
    assert set(ckpt["state_dict"]) == set(m2.state_dict())

    # the same first epoch with the sweeps merged on the device instead of on the loader threads
    _, _, h_dev, _ = run(tmp_path / "c", 2, until=1, dataset=NuScenesDatasetSSL(ann, pipeline=TRAIN_PIPELINE, device_merge=True))
    for a, b in zip(h_dev, h2[:3]):
        assert abs(a - b) <= 2e-3 * abs(b)

    # one epoch, then resume for the second: same schedule position, same optimiser state
    torch.manual_seed(11)
    m1, t1, h1, _ = run(tmp_path / "b", 2, until=1)         # "interrupted" after the first of two epochs
    for a, b in zip(h1, h2[:3]):
        assert abs(a - b) <= 2e-3 * abs(b)
    rng_state = torch.get_rng_state()
    mb = build()
    torch.set_rng_state(rng_state)
    loader = BatchLoader(ds, samples_per_gpu=2, seed=1, workers=2)
    tb, hb = train(mb, loader, str(tmp_path / "b"), 2, base_lr=1e-4, resume_from=str(tmp_path / "b" / "epoch_1.pth"),
                   log_interval=2, log=lambda *_: None)
    assert tb.step_count == 6 and len(hb) == 3
    for a, b in zip(hb, h2[3:]):
        assert abs(a - b) <= 2e-3 * abs(b)       # run-to-run noise of the scatter's float atomics after three updates
    d = (tb.flat_param - t2.flat_param).abs()
    # the cyclic schedule peaks at 100 x base_lr = 1e-2 inside these six iterations; an AdamW update of a noise-level
    # gradient is +-lr, so single weights may differ by a few lr while the bulk agrees
    assert d.max().item() <= 2.5e-2 and d.mean().item() <= 2e-4, (d.max().item(), d.mean().item())
    meta = load_checkpoint(str(tmp_path / "b" / "epoch_2.pth"), build())
    assert meta["epoch"] == 2 and meta["iter"] == 6


def test_train_cli_from_a_config_file(tmp_path):
    """tools/train.py: config file -> dataset -> loader -> runner, two tiny epochs, then resume for a third."""
    import subprocess
    import sys
    ann = write_dataset(tmp_path, n_scenes=4)
    own = os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py")
    cfg = tmp_path / "cfg.py"
    cfg.write_text(f"_base_ = [{own!r}]\n"
                   f"data = dict(samples_per_gpu=2, workers_per_gpu=2, train=dict(data_root={str(tmp_path) + '/'!r}, "
                   f"ann_file={ann!r}))\n"
                   "runner = dict(max_epochs=2)\n")
    work = tmp_path / "work"
    cmd = [sys.executable, os.path.join(ROOT, "tools/train.py"), str(cfg), "--work-dir", str(work), "--impl", "tc1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Epoch [2][2/2]" in out.stdout and sorted(os.listdir(work)) == ["epoch_1.pth", "epoch_2.pth"]
    out = subprocess.run(cmd + ["--max-epochs", "3", "--resume-from", str(work / "epoch_2.pth")], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Epoch [3][2/2]" in out.stdout and "Epoch [1]" not in out.stdout and os.path.exists(work / "epoch_3.pth")
