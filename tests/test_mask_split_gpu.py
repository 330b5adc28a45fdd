"""Device-side random visible/masked split (csrc/mask_split.cu) — the properties the reference's
per-sample randperm split guarantees (…_ssl.py:287-304): exact keep count int(L*(1-ratio)) per frame,
a partition of every frame's pillars, and randomness that follows the seed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def split(counts, keep_frac, seed):
    from geomae_b200 import lib as L
    dev = torch.device("cuda:0")
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    fs = torch.from_numpy(starts).to(dev)
    n_keep = sum(int(n * keep_frac) for n in counts)
    keep = torch.full((n_keep,), -1, dtype=torch.int64, device=dev)
    mask = torch.full((int(starts[-1]) - n_keep,), -1, dtype=torch.int64, device=dev)
    L.run("mask_split", L.ptr(fs), len(counts), float(keep_frac), seed, L.ptr(keep), L.ptr(mask), L.stream_ptr(dev))
    torch.cuda.synchronize()
    return keep.cpu().numpy(), mask.cpu().numpy(), starts


@pytest.mark.parametrize("counts", [[6089, 6120, 5990, 6291], [1], [0, 5, 0, 1500], [3, 2, 1, 0], [85000, 33000]])
def test_partition_and_counts(counts):
    keep_frac = 1 - 0.7
    keep, mask, starts = split(counts, keep_frac, 1234)
    assert np.array_equal(np.sort(np.concatenate([keep, mask])), np.arange(starts[-1]))
    ko = mo = 0
    for f, n in enumerate(counts):
        k = int(n * keep_frac)
        kf, mf = keep[ko:ko + k], mask[mo:mo + n - k]
        assert ((kf >= starts[f]) & (kf < starts[f + 1])).all() and ((mf >= starts[f]) & (mf < starts[f + 1])).all()
        assert (np.diff(kf) > 0).all() and (np.diff(mf) > 0).all()
        ko += k
        mo += n - k


def test_seed_dependence_and_uniformity():
    counts = [20000]
    a, _, _ = split(counts, 0.3, 1)
    b, _, _ = split(counts, 0.3, 1)
    c, _, _ = split(counts, 0.3, 2)
    assert np.array_equal(a, b)
    inter = np.intersect1d(a, c).size
    assert abs(inter - 0.3 * len(a)) < 0.05 * len(a)        # independent subsets overlap by ~30 %
    hist = np.histogram(a, bins=10, range=(0, 20000))[0]   # spread evenly over the index range
    assert hist.min() > 480 and hist.max() < 720


def test_detector_uses_the_kernel_and_follows_manual_seed():
    import geomae_b200  # noqa: F401
    from geomae_b200.registry import Config, build_model
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = Config.fromfile(os.path.join(root, "configs/mae_sst/geomae_nus_pretrain.py"))
    model = build_model(cfg.model)
    coors = torch.zeros((300, 4), dtype=torch.int32, device="cuda:0")
    coors[100:, 0] = 1
    torch.manual_seed(7)
    k1, m1 = model.get_vanilla_mask_index(coors, 2)
    torch.manual_seed(7)
    k2, m2 = model.get_vanilla_mask_index(coors, 2)
    assert torch.equal(k1, k2) and torch.equal(m1, m2)
    assert k1.numel() == int(100 * (1 - 0.7)) + int(200 * (1 - 0.7)) and k1.numel() + m1.numel() == 300
