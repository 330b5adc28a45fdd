"""geomae_b200.ops.DynamicScatter / scatter_v2 — the cases of the reference's own test for this op
(tests/test_models/test_voxel_encoder/test_dynamic_scatter.py:8-93): empty input, empty reduced output, brute-force
mean / max over voxels incl. rows with -1 coordinates, and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ARGS = ([0.32, 0.32, 6], [-74.88, -74.88, -2, 74.88, 74.88, 4])


def _brute(feats, coors):
    ref_coors = coors.unique(dim=0, sorted=True)
    ref_coors = ref_coors[ref_coors.min(dim=-1).values >= 0]
    mean, mx = [], []
    for c in ref_coors:
        sel = feats[(coors == c).all(dim=-1)]
        mean.append(sel.mean(dim=0))
        mx.append(sel.max(dim=0).values)
    return ref_coors, torch.stack(mean), torch.stack(mx)


def test_dynamic_scatter_reference_cases():
    from geomae_b200.ops import DynamicScatter
    dev = "cuda"
    dsmean, dsmax = DynamicScatter(*ARGS, True), DynamicScatter(*ARGS, False)
    # empty input
    ef = torch.empty((0, 3), device=dev, requires_grad=True)
    ec = torch.empty((0, 3), dtype=torch.int32, device=dev)
    for ds in (dsmean, dsmax):
        of, oc = ds(ef, ec)
        of.sum().backward()
        assert of.shape == ef.shape and oc.shape == ec.shape
    # empty reduced output: every row carries a -1
    f = (torch.rand((20000, 3), device=dev) * 100 - 50).requires_grad_()
    c = torch.randint(-1, 0, (20000, 3), dtype=torch.int32, device=dev)
    for ds in (dsmean, dsmax):
        of, oc = ds(f, c)
        assert of.shape[0] == 0 and oc.shape[0] == 0
        of.sum().backward()
        assert (f.grad == 0).all()
    # non-empty input against the brute-force reduction
    feats = torch.rand((20000, 3), device=dev) * 100 - 50
    coors = torch.randint(-1, 8, (20000, 3), dtype=torch.int32, device=dev)
    ref_coors, ref_mean, ref_max = _brute(feats, coors)
    fm, cm = dsmean(feats, coors)
    fx, cx = dsmax(feats, coors)
    assert torch.equal(cm, ref_coors) and torch.equal(cx, ref_coors)
    torch.testing.assert_close(fm, ref_mean, rtol=1e-5, atol=1e-4)
    assert torch.equal(fx, ref_max)
    # batched form: rows ordered by batch first
    bc = torch.cat([torch.randint(0, 3, (20000, 1), dtype=torch.int32, device=dev), coors], dim=1)
    fb, cb = dsmean(feats, bc)
    rb, rm, _ = _brute(feats, bc)
    assert torch.equal(cb, rb)
    torch.testing.assert_close(fb, rm, rtol=1e-5, atol=1e-4)


def test_scatter_v2_modes_and_gradients():
    from geomae_b200.ops import scatter_v2
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    feat = torch.randn((5000, 16), device=dev, generator=g)
    coors = torch.randint(0, 6, (5000, 4), dtype=torch.int32, device=dev, generator=g)
    uniq, inv = torch.unique(coors, return_inverse=True, dim=0)
    for mode in ("sum", "avg", "max"):
        x = feat.clone().requires_grad_()
        y = feat.clone().requires_grad_()
        out, new_coors, unq_inv = scatter_v2(x, coors, mode)
        assert torch.equal(new_coors, uniq) and torch.equal(unq_inv, inv)
        red = {"sum": "sum", "avg": "mean", "max": "amax"}[mode]
        ref = torch.zeros_like(out).scatter_reduce(0, inv[:, None].expand(-1, 16), y, red, include_self=False)
        torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        (ref * w).sum().backward()
        torch.testing.assert_close(x.grad, y.grad, rtol=1e-5, atol=1e-5)
    # min_points drops sparse voxels; a given unq_inv / new_coors pair is honoured
    out, nc = scatter_v2(feat, coors, "max", return_inv=False, min_points=5)
    cnt = torch.bincount(inv)
    assert torch.equal(nc, uniq[cnt >= 5])
    out2, nc2, inv2 = scatter_v2(feat, coors, "sum", unq_inv=inv, new_coors=uniq)
    assert torch.equal(nc2, uniq) and torch.equal(inv2, inv)


def test_unique_rows_matches_torch_unique():
    """scatter_v2's row ranking (bitmap + prefix sum, geomae_coors_rank) == torch.unique(dim=0): same unique rows in
    the same (lexicographic) order, same inverse map, same counts — for (z,y,x) and (b,z,y,x) rows with duplicates,
    sub-voxel sized grids, a single row, and the fallback for negative rows."""
    from geomae_b200.ops import unique_rows
    g = torch.Generator().manual_seed(0)
    cases = [torch.stack([torch.randint(0, hi, (n,), generator=g) for hi in his], dim=1)
             for n, his in ((5000, (3, 8, 200, 200)), (20000, (8, 1600, 1600)), (1, (2, 1, 40, 40)),
                            (3000, (4, 1, 400, 400)), (4000, (2, 4, 30, 17)))]
    cases.append(torch.tensor([[0, 0, 5, 7]] * 9 + [[1, 0, 0, 0]]))
    for dtype in (torch.int32, torch.int64):
        for c in cases:
            c = c.to(DEV).to(dtype)
            uniq, inv, cnt = unique_rows(c)
            ru, ri, rc = torch.unique(c, return_inverse=True, return_counts=True, dim=0)
            assert uniq.dtype == c.dtype and torch.equal(uniq, ru)
            assert torch.equal(inv, ri) and torch.equal(cnt, rc)
    neg = torch.tensor([[0, -1, -1, -1], [0, 0, 2, 3], [0, 0, 2, 3]], device=DEV)
    uniq, inv, cnt = unique_rows(neg)
    assert torch.equal(uniq, torch.unique(neg, dim=0)) and cnt.tolist() == [1, 2]
