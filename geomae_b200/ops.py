"""The reference's Python op surface for this path — ``from mmdet3d.ops import Voxelization, DynamicScatter,
dynamic_scatter, scatter_v2, flat2window, window2flat, get_flat2win_inds, get_inner_win_inds,
make_continuous_inds`` (mmdet3d/ops/__init__.py:22-26) — with the same names, argument meaning and return
conventions, so code written against the reference (its model files, its tests' idiom) can call into this library.

The reductions run in the sm_100a kernels behind the C ABI (geomae_scatter_reduce_fwd/bwd); there is no CPU
fallback: CPU tensors raise.  The padded-window helpers (flat2window & co.) are index bookkeeping the production
path no longer needs — it attends over CSR windows (windows.py, csrc/window_csr.cu) — and are kept as plain tensor
ops for interface compatibility and for checking the CSR layout against the reference's bucketed one.
"""
from __future__ import annotations

import torch
from torch import nn

from . import lib as L
from .voxel import Voxelization  # noqa: F401  (re-export: mmdet3d/ops/voxel/voxelize.py:63-112)

_MODES = {"sum": 0, "mean": 1, "avg": 1, "max": 2}


class _ScatterRows(torch.autograd.Function):
    """new_feat[v] = reduce over {p : inv[p] == v} feat[p]  (torch_scatter.scatter / scatter_max of sst_ops.py:29-32,
    and dynamic_point_to_voxel_forward/backward of mmdet3d/ops/voxel/src/voxelization.h:122-154)."""

    @staticmethod
    def forward(ctx, feat, inv, n_out, mode):
        L.require_cuda(feat, "feat")
        feat = feat.contiguous().float()
        n, c = feat.shape
        inv32 = inv.to(torch.int32).contiguous()
        mean = torch.zeros((max(n_out, 1), 4), dtype=torch.float32, device=feat.device)
        if mode == 1:       # counts live in column 3 of the [V,4] per-voxel record the kernels read
            mean[:, 3] = torch.bincount(inv, minlength=max(n_out, 1)).float()
        out = torch.empty((n_out, c), dtype=torch.float32, device=feat.device)
        arg = torch.empty((n_out, c), dtype=torch.int32, device=feat.device) if mode == 2 else None
        L.run("scatter_reduce_fwd", L.ptr(feat), n, c, L.ptr(inv32), L.ptr(mean), n_out, mode, L.ptr(out), L.ptr(arg),
              L.stream_ptr(feat.device))
        ctx.mode, ctx.shape = mode, (n, c)
        ctx.save_for_backward(inv32, mean, arg)
        return out

    @staticmethod
    def backward(ctx, d_out):
        inv32, mean, arg = ctx.saved_tensors
        n, c = ctx.shape
        d_feat = torch.empty((n, c), dtype=torch.float32, device=d_out.device)
        L.run("scatter_reduce_bwd", L.ptr(d_out.contiguous()), n, c, L.ptr(inv32), L.ptr(mean), L.ptr(arg), ctx.mode,
              L.ptr(d_feat), L.stream_ptr(d_out.device))
        return d_feat, None, None, None


_MAX_CELLS = (1 << 31) - 1


def unique_rows(coors):
    """``torch.unique(coors, return_inverse=True, return_counts=True, dim=0)`` for non-negative integer grid rows
    [N, 3] (z,y,x) or [N, 4] (b,z,y,x) without the library's lexicographic sort: the rows set bits in an occupancy
    bitmap over the bounding grid, a prefix sum ranks the set bits (= the sorted unique rows), every row looks its rank
    up (geomae_coors_rank, csrc/voxel_scatter.cu).  Falls back to torch.unique only when the bounding grid has 2^31
    cells or more, or a coordinate is negative (both on the device; the reference's -1 rows are filtered upstream)."""
    import ctypes as C
    from .voxel import VoxelGeometry
    n, w = coors.shape
    if n == 0 or w not in (3, 4) or not coors.is_cuda:
        return torch.unique(coors, return_inverse=True, return_counts=True, dim=0)
    lo = coors.min().item()
    ext = (coors.max(dim=0)[0] + 1).tolist()
    b_ext, (gz, gy, gx) = (ext[0], ext[1:]) if w == 4 else (1, ext)
    if lo < 0 or b_ext * gz * gy * gx > _MAX_CELLS:
        return torch.unique(coors, return_inverse=True, return_counts=True, dim=0)
    dev = coors.device
    c32 = coors.to(torch.int32)
    folded = torch.zeros((n, 4), dtype=torch.int32, device=dev)          # (b, 0, z*Y + y, x)
    if w == 4:
        folded[:, 0] = c32[:, 0]
    folded[:, 2] = c32[:, -3] * gy + c32[:, -2]
    folded[:, 3] = c32[:, -1]
    geom = VoxelGeometry((0.0, 0.0, 0.0, float(gx), float(gz * gy), 1.0), (1.0, 1.0, 1.0), (1.0, 1.0, 1.0),
                         (1.0, 1.0, 1.0), (1, 1, 1), (1, 1, 1))
    n_words = (b_ext * gz * gy * gx + 31) // 32
    i32 = dict(dtype=torch.int32, device=dev)
    bitmap, word_rank = torch.empty(n_words, **i32), torch.empty(n_words, **i32)
    scan_tmp, counts = torch.empty(3 * 16384, **i32), torch.empty(4, **i32)
    rank, first = torch.empty(n, **i32), torch.empty(n, **i32)
    L.run("coors_rank", C.byref(geom.cstruct), L.ptr(folded), n, b_ext, L.ptr(bitmap), L.ptr(word_rank), L.ptr(scan_tmp),
          L.ptr(counts), L.ptr(rank), L.ptr(first), L.stream_ptr(dev))
    n_unique = int(counts[0].item())
    inv = rank.long()
    new_coors = coors.index_select(0, first[:n_unique].long())
    return new_coors, inv, torch.bincount(inv, minlength=n_unique)


def scatter_v2(feat, coors, mode, return_inv=True, min_points=0, unq_inv=None, new_coors=None):
    """mmdet3d/ops/sst/sst_ops.py:8-39.  ``new_coors`` are the unique rows of ``coors`` in lexicographic order,
    ``unq_inv`` maps every point to its row, ``new_feat`` is the per-row 'sum' | 'mean' ('avg') | 'max' of ``feat``."""
    assert feat.size(0) == coors.size(0)
    if mode not in _MODES:
        raise NotImplementedError(mode)
    counts = None
    if unq_inv is None:
        new_coors, unq_inv, counts = unique_rows(coors)
    else:
        assert new_coors is not None, "please pass new_coors for interface consistency"
    if min_points > 0:
        if counts is None:
            counts = torch.bincount(unq_inv, minlength=new_coors.shape[0])
        valid = counts[unq_inv] >= min_points
        feat, coors = feat[valid], coors[valid]
        new_coors, unq_inv, _ = unique_rows(coors)
    new_feat = _ScatterRows.apply(feat, unq_inv, new_coors.shape[0], _MODES[mode])
    return (new_feat, new_coors, unq_inv) if return_inv else (new_feat, new_coors)


def dynamic_scatter(feats, coors, reduce_type="max"):
    """mmdet3d/ops/voxel/scatter_points.py:9-45: (voxel_feats, voxel_coors); rows with a negative coordinate are
    dropped (the -1 rows upstream dynamic voxelisation emits for out-of-range points)."""
    keep = (coors >= 0).all(dim=1)
    if not bool(keep.all()):
        feats, coors = feats[keep], coors[keep]
    voxel_feats, voxel_coors = scatter_v2(feats, coors, reduce_type, return_inv=False)
    return voxel_feats, voxel_coors


class DynamicScatter(nn.Module):
    """mmdet3d/ops/voxel/scatter_points.py:53-109.  ``coors`` is [N,3] (z,y,x) or [N,4] with a leading batch index;
    the batched form is reduced in ONE call (rows come out ordered by batch first, exactly the reference's
    per-sample loop + concatenation)."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors):
        return dynamic_scatter(points.contiguous(), coors.contiguous(), "mean" if self.average_points else "max")

    def forward(self, points, coors):
        return self.forward_single(points, coors)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range={self.point_cloud_range}, "
                f"average_points={self.average_points})")


# ------------------------------------------------------------------------------------------------ padded-window helpers
@torch.no_grad()
def make_continuous_inds(inds):
    """sst_ops.py:371-388: replace arbitrary non-negative ids by their rank among the distinct ids (0 .. K-1)."""
    return torch.unique(inds, sorted=True, return_inverse=True)[1].to(inds.dtype)


@torch.no_grad()
def get_inner_win_inds(win_inds):
    """sst_ops.py:271-319: for every element its position 0 .. m-1 among the m elements carrying the same id.  The
    reference's order inside a window follows an unstable sort, i.e. is unspecified; here it is the original order."""
    n = win_inds.shape[0]
    if n == 0:
        return win_inds.clone()
    order = torch.argsort(win_inds, stable=True)
    sorted_ids = win_inds[order]
    pos = torch.arange(n, device=win_inds.device, dtype=win_inds.dtype)
    is_start = torch.ones(n, dtype=torch.bool, device=win_inds.device)
    is_start[1:] = sorted_ids[1:] != sorted_ids[:-1]
    start = torch.cummax(torch.where(is_start, pos, torch.zeros_like(pos)), dim=0)[0]
    inner = torch.empty_like(win_inds)
    inner[order] = pos - start
    return inner


@torch.no_grad()
def get_flat2win_inds(batch_win_inds, voxel_drop_lvl, drop_info, debug=True):
    """sst_ops.py:57-95: per drop level dl -> (flat2window_inds, (positions,)) with
    flat2window_inds = rank(window) * max_tokens + position inside the window."""
    out = {}
    for dl in drop_info:
        mask = voxel_drop_lvl == dl
        if not bool(mask.any()):
            continue
        win = make_continuous_inds(batch_win_inds[mask])
        max_tokens = drop_info[dl]["max_tokens"]
        inner = get_inner_win_inds(win)
        if debug:
            assert int(inner.max()) < max_tokens, f"a window holds more than max_tokens={max_tokens} voxels"
        out[dl] = (win * max_tokens + inner, torch.where(mask))
    return out


def flat2window(feat, voxel_drop_lvl, flat2win_inds_dict, drop_info):
    """sst_ops.py:98-135: {dl: zero-padded [num_windows, max_tokens, C]} from flat [N, C] rows."""
    out = {}
    for dl in drop_info:
        mask = voxel_drop_lvl == dl
        if not bool(mask.any()):
            continue
        inds = flat2win_inds_dict[dl][0]
        max_tokens = drop_info[dl]["max_tokens"]
        n_win = int(torch.div(inds, max_tokens, rounding_mode="floor").max()) + 1
        padded = feat.new_zeros((n_win * max_tokens, feat.shape[-1]))
        padded = padded.index_copy(0, inds, feat[mask])
        out[dl] = padded.reshape(n_win, max_tokens, feat.shape[-1])
    return out


def window2flat(feat_3d_dict, inds_dict):
    """sst_ops.py:225-251: inverse of flat2window."""
    first = next(iter(feat_3d_dict.values()))
    n = sum(v[0].shape[0] for v in inds_dict.values())
    flat = first.new_zeros((n, first.shape[-1]))
    seen = torch.zeros(n, dtype=torch.bool, device=first.device)
    for dl, feat in feat_3d_dict.items():
        inds, flat_pos = inds_dict[dl]
        rows = feat.reshape(-1, feat.shape[-1]).index_select(0, inds)
        flat = flat.index_copy(0, flat_pos[0], rows)
        seen[flat_pos[0]] = True
    assert bool(seen.all()), "window2flat: some voxels are not covered by inds_dict"
    return flat
