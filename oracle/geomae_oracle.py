"""TEST INFRASTRUCTURE — CPU restatement of GeoMAE's masked-pretraining hot path.

This file is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``geomae_b200``) never does; it fails loudly without its CUDA
library instead of falling back to anything here.

Parity pin: every function below is checked against the *unmodified reference*
executed on CPU by ``oracle/ref_harness.py`` (see ``oracle/make_golden.py`` and
``tests/test_oracle_vs_golden.py``); the resulting vectors are committed under
``tests/golden/``.  The reference's own tests pin only dynamic voxelisation and
scatter mean/max; everything else is pinned by those goldens.

Integer/index stages are numpy; floating-point model stages are torch-CPU fp32
(the same library arithmetic the reference itself uses).  Paths in the
docstrings are relative to the reference tree.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

CURV_EPS = 1e-9  # detectors/multi_sub_voxel_dynamic_voxelnet_ssl.py:19


@dataclass
class PathConfig:
    """Constants of configs/mae_sst/m_sst_nus_singlestage_curv_07_ssl_dataset_wo_dbsampler_6x_1e-5.py."""
    pc_range: tuple = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)          # :16
    voxel_size: tuple = (0.256, 0.256, 8)                            # :14 (x,y,z)
    sub_voxel_size_med: tuple = (0.128, 0.128, 2)                    # :19
    sub_voxel_size_low: tuple = (0.064, 0.064, 1)                    # :18
    sub_voxel_ratio_med: tuple = (4, 2, 2)                           # :22 (z,y,x)
    sub_voxel_ratio_low: tuple = (8, 4, 4)                           # :21 (z,y,x)
    grid_size: tuple = (1, 400, 400)                                 # :24 (z,y,x)
    window_shape: tuple = (12, 12)                                   # :15
    shifts: tuple = ((0, 0), (6, 6))                                 # :57
    drop_info: dict = field(default_factory=lambda: {                # :38-41 (training)
        0: dict(max_tokens=56, drop_range=(0, 56)),
        1: dict(max_tokens=144, drop_range=(56, 100000))})
    mask_ratio: float = 0.7                                          # :25
    loss_low: float = 10.0                                           # :26
    loss_med: float = 8.0                                            # :27
    loss_top: float = 10.0                                           # :28
    loss_nor: float = 4.0                                            # :33
    cls_low: float = 5.0                                             # :29
    cls_med: float = 2.0                                             # :30
    d_model: int = 128                                               # :139
    nhead: int = 8                                                   # :140
    ffn: int = 256                                                   # :145
    enc_blocks: int = 6                                              # :143
    dec_blocks: int = 2                                              # :144
    bn_eps: float = 1e-3                                             # :124
    bn_momentum: float = 0.01                                        # :124
    pos_temperature: float = 10000.0                                 # :157

    @property
    def slots_med(self):
        return int(np.prod(self.sub_voxel_ratio_med))

    @property
    def slots_low(self):
        return int(np.prod(self.sub_voxel_ratio_low))


# ---------------------------------------------------------------------------
# a1/a2  dynamic voxelisation
# ---------------------------------------------------------------------------
def grid_shape_xyz(voxel_size, pc_range):
    """ops/voxel/src/voxelization_cpu.cpp:153-156: ceil of the fp32 quotient."""
    lo = np.asarray(pc_range[:3], np.float32)
    hi = np.asarray(pc_range[3:], np.float32)
    vs = np.asarray(voxel_size, np.float32)
    return np.ceil((hi - lo) / vs).astype(np.int32)


def dynamic_voxelize(points: np.ndarray, voxel_size, pc_range) -> np.ndarray:
    """ops/voxel/src/voxelization_cpu.cpp:6-40 — fp32 subtract, IEEE divide, floor,
    clamp into [0, grid-1] (this fork clamps instead of writing -1), stored (z,y,x)."""
    pts = np.asarray(points, np.float32)
    lo = np.asarray(pc_range[:3], np.float32)
    vs = np.asarray(voxel_size, np.float32)
    grid = grid_shape_xyz(voxel_size, pc_range)
    c = np.floor((pts[:, :3] - lo) / vs).astype(np.int32)
    c = np.minimum(np.maximum(c, 0), grid - 1)
    return np.ascontiguousarray(c[:, ::-1])


def batch_voxelize(frames, voxel_size, pc_range) -> np.ndarray:
    """detectors/…_ssl.py:307-377 — per-sample voxelise, prepend batch index -> [P,4] (b,z,y,x)."""
    out = []
    for b, pts in enumerate(frames):
        c = dynamic_voxelize(pts, voxel_size, pc_range)
        out.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
    return np.concatenate(out, axis=0)


def unique_rows(coors: np.ndarray):
    """torch.unique(dim=0, return_inverse, return_counts) on (b,z,y,x) rows
    (ops/sst/sst_ops.py:15-17, detectors/…_ssl.py:749): lexicographically sorted rows."""
    c = coors.astype(np.int64)
    span = c.max(axis=0) + 1
    key = ((c[:, 0] * span[1] + c[:, 1]) * span[2] + c[:, 2]) * span[3] + c[:, 3]
    uniq, first, inv, cnt = np.unique(key, return_index=True, return_inverse=True, return_counts=True)
    return coors[first], inv.astype(np.int64), cnt.astype(np.int64)


# ---------------------------------------------------------------------------
# a6/a7/a8/a10/a11  centroids, slot tensors, neighbour table
# ---------------------------------------------------------------------------
def centroid_per_voxel(xyz_zyx: np.ndarray, coors: np.ndarray):
    """detectors/…_ssl.py:726-768 — fp32 scatter-add (sequential point order on CPU) / count."""
    rows, inv, cnt = unique_rows(coors)
    acc = np.zeros((rows.shape[0], 3), np.float32)
    np.add.at(acc, inv, xyz_zyx.astype(np.float32))
    return acc / cnt.astype(np.float32)[:, None], rows, cnt


def pillar_lookup(pillar_coors: np.ndarray, batch_size: int, grid_size):
    """Dense pillar->row table (detectors/…_ssl.py:654-658): key b*Z*Y*X + y*Y + x.
    Unset cells stay 0 exactly like the reference's new_zeros table."""
    gz, gy, gx = grid_size
    table = np.zeros(batch_size * gz * gy * gx, np.int64)
    pc = pillar_coors.astype(np.int64)
    table[pc[:, 0] * gz * gy * gx + pc[:, 2] * gy + pc[:, 3]] = np.arange(pc.shape[0])
    return table


def sub_voxel_slots(pillar_coors, sub_coors, ratio, batch_size, grid_size):
    """detectors/…_ssl.py:659-665 / :696-702 — parent pillar row and slot id of every sub-voxel."""
    gz, gy, gx = grid_size
    table = pillar_lookup(pillar_coors, batch_size, grid_size)
    sc = sub_coors.astype(np.int64)
    parent = table[sc[:, 0] * gz * gy * gx + (sc[:, 2] // ratio[1]) * gy + sc[:, 3] // ratio[2]]
    slot = (sc[:, 1] % ratio[0]) * (ratio[1] * ratio[2]) + (sc[:, 2] % ratio[1]) * ratio[2] + sc[:, 3] % ratio[2]
    return parent, slot


def dense_slots(n_pillars, parent, slot, values, n_slots):
    """Scatter per-sub-voxel rows into [V, n_slots, 3] + bool mask (…_ssl.py:650-669)."""
    dense = np.zeros((n_pillars * n_slots, 3), np.float32)
    mask = np.zeros(n_pillars * n_slots, bool)
    flat = parent * n_slots + slot
    dense[flat] = values
    mask[flat] = True
    return dense.reshape(n_pillars, n_slots, 3), mask.reshape(n_pillars, n_slots)


def neighbour_pairs(pillar_coors, batch_size, grid_size):
    """spconv 2.1.21 get_indice_pairs_implicit_gemm(subm, ksize=[1,3,3]) as called at
    detectors/…_ssl.py:192-207, restated: pair[(dy+1)*3+(dx+1), i] = row of the pillar at
    (y+dy, x+dx) in the same sample, -1 if empty/outside."""
    _, gy, gx = grid_size
    pc = pillar_coors.astype(np.int64)
    n = pc.shape[0]
    table = np.full(batch_size * gy * gx, -1, np.int64)
    table[pc[:, 0] * gy * gx + pc[:, 2] * gx + pc[:, 3]] = np.arange(n)
    pair = np.full((9, n), -1, np.int32)
    k = 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            y, x = pc[:, 2] + dy, pc[:, 3] + dx
            ok = (y >= 0) & (y < gy) & (x >= 0) & (x < gx)
            key = pc[:, 0] * gy * gx + np.clip(y, 0, gy - 1) * gx + np.clip(x, 0, gx - 1)
            pair[k] = np.where(ok, table[key], -1)
            k += 1
    return pair


def scatter_matrix(med_dense, med_mask, centroid_top, pair):
    """detectors/…_ssl.py:583-597 — X = neighbours' med centroids minus the centre pillar's
    centroid over 9x16 slots (absent slots exactly zero), C = X^T X (not divided by count)."""
    v, s, _ = med_dense.shape
    safe = np.where(pair < 0, 0, pair)
    around = med_dense[safe]                      # [9,V,16,3]
    amask = med_mask[safe] & (pair >= 0)[:, :, None]
    around = np.where(amask[..., None], around, np.float32(0))
    around = around.transpose(1, 0, 2, 3).reshape(v, 9 * s, 3)
    amask = amask.transpose(1, 0, 2).reshape(v, 9 * s)
    centre = np.where(amask[..., None], centroid_top[:, None, :], np.float32(0))
    x = torch.from_numpy(np.ascontiguousarray(around - centre))
    return (x.transpose(-2, -1) @ x), amask.sum(axis=1)


def normal_and_curvature(cov: torch.Tensor):
    """detectors/…_ssl.py:598-607 — torch.svd; normal = last right-singular vector,
    re-normalised; curvature = (S + 1e-9) / sum in float64.  The normal's SIGN is a LAPACK
    artefact; callers compare up to sign."""
    u, s, v = torch.svd(cov)
    normal = v[..., -1]
    normal = normal / torch.norm(normal, p=2, dim=-1, keepdim=True)
    curv = s.to(torch.float64) + CURV_EPS
    curv = curv / curv.sum(dim=-1, keepdim=True)
    return normal, curv, s


def normalize_centroids(coors_zyx, centroids, voxel_size_xyz, pc_range):
    """detectors/…_ssl.py:626-641 — (c - (coor*size + min)) / size with sizes reversed to (z,y,x)."""
    vs = torch.tensor(tuple(voxel_size_xyz)[::-1], dtype=torch.float32)
    lo = torch.tensor(tuple(pc_range[:3])[::-1], dtype=torch.float32)
    corner = torch.from_numpy(coors_zyx.astype(np.int64)) * vs + lo
    return ((torch.from_numpy(centroids) - corner) / vs).numpy()


def geometric_targets(frames, cfg: PathConfig, ids_mask: np.ndarray | None = None):
    """Everything target-side of extract_feat (detectors/…_ssl.py:169-235) except the model."""
    b = len(frames)
    pts = np.concatenate(frames, axis=0).astype(np.float32)
    xyz_zyx = pts[:, [2, 1, 0]]
    coors_top = batch_voxelize(frames, cfg.voxel_size, cfg.pc_range)
    coors_med = batch_voxelize(frames, cfg.sub_voxel_size_med, cfg.pc_range)
    coors_low = batch_voxelize(frames, cfg.sub_voxel_size_low, cfg.pc_range)
    cen_low, rows_low, cnt_low = centroid_per_voxel(xyz_zyx, coors_low)
    cen_med, rows_med, cnt_med = centroid_per_voxel(xyz_zyx, coors_med)
    cen_top, rows_top, cnt_top = centroid_per_voxel(xyz_zyx, coors_top)
    v = rows_top.shape[0]
    par_m, slot_m = sub_voxel_slots(rows_top, rows_med, cfg.sub_voxel_ratio_med, b, cfg.grid_size)
    par_l, slot_l = sub_voxel_slots(rows_top, rows_low, cfg.sub_voxel_ratio_low, b, cfg.grid_size)
    med_raw, med_mask = dense_slots(v, par_m, slot_m, cen_med, cfg.slots_med)
    pair = neighbour_pairs(rows_top, b, cfg.grid_size)
    cov, n_contrib = scatter_matrix(med_raw, med_mask, cen_top, pair)
    normal, curv, sing = normal_and_curvature(cov)
    n_low = normalize_centroids(rows_low[:, 1:], cen_low, cfg.sub_voxel_size_low, cfg.pc_range)
    n_med = normalize_centroids(rows_med[:, 1:], cen_med, cfg.sub_voxel_size_med, cfg.pc_range)
    n_top = normalize_centroids(rows_top[:, 1:], cen_top, cfg.voxel_size, cfg.pc_range)
    low_dense, low_mask = dense_slots(v, par_l, slot_l, n_low, cfg.slots_low)
    med_dense, med_mask2 = dense_slots(v, par_m, slot_m, n_med, cfg.slots_med)
    out = dict(coors_top=coors_top, coors_med=coors_med, coors_low=coors_low,
               pillar_coors=rows_top, pillar_count=cnt_top, centroid_top=cen_top,
               rows_med=rows_med, centroid_med=cen_med, rows_low=rows_low, centroid_low=cen_low,
               med_raw=med_raw, med_mask=med_mask, pair=pair, cov=cov.numpy(), n_contrib=n_contrib,
               normal=normal.numpy(), curvature=curv.numpy(), singular=sing.numpy(),
               norm_top=n_top, low_dense=low_dense, low_mask=low_mask,
               med_dense=med_dense)
    if ids_mask is not None:
        m = np.asarray(ids_mask)
        out.update(tgt_low=low_dense[m], tgt_low_mask=low_mask[m], tgt_med=med_dense[m],
                   tgt_med_mask=med_mask2[m], tgt_top=n_top[m], tgt_normal=out["normal"][m],
                   tgt_curv=out["curvature"][m], mask_coors=rows_top[m])
    return out


def vanilla_mask_ids(pillar_coors, batch_size, ratio, seed):
    """detectors/…_ssl.py:287-304 with a seeded CPU generator: per sample randperm(L),
    keep the first int(L*(1-ratio)) (Python float64 arithmetic), mask the rest."""
    g = torch.Generator().manual_seed(seed)
    keep, mask = [], []
    pc = torch.from_numpy(np.asarray(pillar_coors))
    for b in range(batch_size):
        inds = torch.where(pc[:, 0] == b)[0]
        n = inds.shape[0]
        len_keep = int(n * (1 - ratio))
        perm = torch.randperm(n, generator=g)
        keep.append(inds[perm[:len_keep]])
        mask.append(inds[perm[len_keep:]])
    return torch.cat(keep).numpy(), torch.cat(mask).numpy()


# ---------------------------------------------------------------------------
# a3/a4  DynamicScatterVFE
# ---------------------------------------------------------------------------
class _ScatterMaxFn(torch.autograd.Function):
    """torch_scatter.scatter_max restated (ops/sst/sst_ops.py:30): per-row max; the gradient
    goes to ONE arg-max point — the smallest point index among ties, the rule of the in-repo
    op (ops/voxel/src/scatter_points_cuda.cu:154-158)."""

    @staticmethod
    def forward(ctx, src, index, n):
        idx = index.view(-1, 1).expand_as(src)
        out = torch.full((n, src.shape[1]), float("-inf"), dtype=src.dtype)
        out = out.scatter_reduce(0, idx, src, "amax", include_self=True)
        rows = torch.arange(src.shape[0]).view(-1, 1).expand_as(src)
        cand = torch.where(src == out[index], rows, torch.full_like(rows, src.shape[0]))
        arg = torch.full((n, src.shape[1]), src.shape[0], dtype=torch.long)
        arg = arg.scatter_reduce(0, idx, cand, "amin", include_self=True)
        ctx.save_for_backward(arg)
        ctx.n_src = src.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        out = g.new_zeros((ctx.n_src, g.shape[1]))
        out.scatter_(0, arg, g)
        return out, None, None


def batch_norm_train(x, weight, bias, eps, sync=None):
    """ops/norm.py:55-86.  Single process: nn.BatchNorm1d training statistics (:58-59).
    ``sync`` = list of the other ranks' [mean, meansqr] vectors emulates the all-gather branch
    (:65-83): equal-weight rank average, var = E[x^2] - E[x]^2."""
    if sync is None:
        return F.batch_norm(x, None, None, weight, bias, True, 0.0, eps)
    mean = x.mean(dim=0)
    meansqr = (x * x).mean(dim=0)
    vec = torch.cat([mean, meansqr])
    vec = (vec + sum(sync)) * (1.0 / (len(sync) + 1))
    mean, meansqr = torch.split(vec, x.shape[1])
    var = meansqr - mean * mean
    scale = weight * torch.rsqrt(var + eps)
    return x * scale.view(1, -1) + (bias - mean * scale).view(1, -1)


def vfe_forward(params, points: torch.Tensor, coors_top: np.ndarray, cfg: PathConfig, prefix="voxel_encoder."):
    """voxel_encoders/voxel_encoder.py:358-419 + utils.py:129-144."""
    rows, inv, _ = unique_rows(coors_top)
    inv_t = torch.from_numpy(inv)
    n = rows.shape[0]
    xyz = points[:, :3]
    cnt = torch.zeros(n).index_add_(0, inv_t, torch.ones(points.shape[0]))
    mean = torch.zeros(n, 3).index_add_(0, inv_t, xyz) / cnt.view(-1, 1)
    f_cluster = xyz - mean[inv_t]
    vx, vy, vz = cfg.voxel_size
    x_off, y_off, z_off = vx / 2 + cfg.pc_range[0], vy / 2 + cfg.pc_range[1], vz / 2 + cfg.pc_range[2]
    c = torch.from_numpy(coors_top)
    f_center = torch.stack([points[:, 0] - (c[:, 3].float() * vx + x_off),
                            points[:, 1] - (c[:, 2].float() * vy + y_off),
                            points[:, 2] - (c[:, 1].float() * vz + z_off)], dim=1)
    feats = torch.cat([points, f_cluster, f_center], dim=1)
    n_layers = 2
    voxel_feats = None
    for i in range(n_layers):
        w = params[f"{prefix}vfe_layers.{i}.linear.weight"]
        x = F.linear(feats, w)
        x = batch_norm_train(x, params[f"{prefix}vfe_layers.{i}.norm.weight"],
                             params[f"{prefix}vfe_layers.{i}.norm.bias"], cfg.bn_eps)
        pf = F.relu(x)
        voxel_feats = _ScatterMaxFn.apply(pf, inv_t, n)
        if i != n_layers - 1:
            feats = torch.cat([pf, voxel_feats[inv_t]], dim=1)
    return voxel_feats, rows, inv


# ---------------------------------------------------------------------------
# a12-a17  window partition / bucketing / position embedding
# ---------------------------------------------------------------------------
def window_partition(coors: np.ndarray, cfg: PathConfig, shift_id: int):
    """backbones/multi_mae_sst_spearate_top_only.py:628-659."""
    wx, wy = cfg.window_shape
    gx, gy = grid_shape_xyz(cfg.voxel_size, cfg.pc_range)[:2]
    nwx = int(np.ceil(gx / wx) + 1)
    nwy = int(np.ceil(gy / wy) + 1)
    sx, sy = cfg.shifts[shift_id]
    c = coors.astype(np.int64)
    x = c[:, 3] + (wx - sx if sx > 0 else 0)
    y = c[:, 2] + (wy - sy if sy > 0 else 0)
    win = c[:, 0] * nwx * nwy + (x // wx) * nwy + (y // wy)
    return win, np.stack([x % wx, y % wy], axis=1)


def window_levels(win: np.ndarray, cfg: PathConfig):
    """backbones/…top_only.py:519-541 — bucket by tokens-per-window, lower <= n < upper."""
    cnt = np.bincount(win)[win]
    lvl = np.full(win.shape, -1, np.int64)
    for dl, info in cfg.drop_info.items():
        lo, hi = info["drop_range"]
        lvl[(cnt >= lo) & (cnt < hi)] = dl
    assert (lvl >= 0).all()
    return lvl, cnt


def flat2win_indices(win: np.ndarray, lvl: np.ndarray, cfg: PathConfig):
    """backbones/…top_only.py:413-507,661-681 — per bucket: window rank (sorted unique),
    position inside the window, flat slot = rank*max_tokens + position.  The reference's
    position comes from an unstable sort; any permutation inside a window is equivalent for
    attention, we use ascending token index."""
    out = {}
    for dl, info in cfg.drop_info.items():
        sel = np.where(lvl == dl)[0]
        if sel.size == 0:
            continue
        uniq, rank = np.unique(win[sel], return_inverse=True)
        order = np.argsort(rank, kind="stable")
        start = np.searchsorted(rank[order], np.arange(uniq.size))
        inner = np.empty(sel.size, np.int64)
        inner[order] = np.arange(sel.size) - start[rank[order]]
        assert inner.max() < info["max_tokens"]
        out[dl] = (rank * info["max_tokens"] + inner, sel, uniq.size)
    return out


def pos_embed_table(cfg: PathConfig) -> torch.Tensor:
    """backbones/…top_only.py:361-394 for every in-window coordinate: [wx*wy, d_model],
    row = cx*wy + cy."""
    wx, wy = cfg.window_shape
    half = cfg.d_model // 2
    inv_freq = torch.arange(half, dtype=torch.float32)
    inv_freq = cfg.pos_temperature ** (2 * (inv_freq // 2) / half)
    cx, cy = torch.meshgrid(torch.arange(wx), torch.arange(wy), indexing="ij")
    x = cx.reshape(-1) - wx / 2
    y = cy.reshape(-1) - wy / 2
    ex = x[:, None] / inv_freq[None, :]
    ey = y[:, None] / inv_freq[None, :]
    ex = torch.stack([ex[:, ::2].sin(), ex[:, 1::2].cos()], dim=-1).flatten(1)
    ey = torch.stack([ey[:, ::2].sin(), ey[:, 1::2].cos()], dim=-1).flatten(1)
    return torch.cat([ex, ey], dim=-1)


class WindowLayout:
    """get_voxel_info (backbones/…top_only.py:143-196) for one token set: both shifts."""

    def __init__(self, coors: np.ndarray, cfg: PathConfig, levels=None):
        """levels: per-shift drop levels decided elsewhere (SSTInputLayer, after its voxel drop) instead of the
        backbone's own bucketing."""
        self.cfg, self.n = cfg, coors.shape[0]
        table = pos_embed_table(cfg)
        self.shifts = []
        for s in range(len(cfg.shifts)):
            win, ciw = window_partition(coors, cfg, s)
            lvl, cnt = window_levels(win, cfg) if levels is None else (levels[s], np.bincount(win)[win])
            inds = flat2win_indices(win, lvl, cfg)
            pos_flat = table[torch.from_numpy(ciw[:, 0] * cfg.window_shape[1] + ciw[:, 1])]
            self.shifts.append(dict(win=win, ciw=ciw, lvl=lvl, cnt=cnt, inds=inds, pos=pos_flat))

    def to_windows(self, shift, feat):
        """ops/sst/sst_ops.py:98-135 flat2window."""
        out = {}
        for dl, (slot, sel, n_win) in self.shifts[shift]["inds"].items():
            t = self.cfg.drop_info[dl]["max_tokens"]
            buf = feat.new_zeros((n_win * t, feat.shape[-1]))
            buf = buf.index_put((torch.from_numpy(slot),), feat[torch.from_numpy(sel)])
            out[dl] = buf.view(n_win, t, -1)
        return out

    def to_flat(self, shift, win_feats):
        """ops/sst/sst_ops.py:225-251 window2flat."""
        c = next(iter(win_feats.values())).shape[-1]
        flat = next(iter(win_feats.values())).new_zeros((self.n, c))
        for dl, (slot, sel, _) in self.shifts[shift]["inds"].items():
            flat = flat.index_put((torch.from_numpy(sel),),
                                  win_feats[dl].reshape(-1, c)[torch.from_numpy(slot)])
        return flat

    def key_padding(self, shift):
        """backbones/…top_only.py:306-316 — True at padded slots."""
        ones = torch.ones((self.n, 1))
        return {dl: v.squeeze(2) == 0 for dl, v in self.to_windows(shift, ones).items()}


# ---------------------------------------------------------------------------
# a18-a21  SRA blocks, backbone
# ---------------------------------------------------------------------------
def sra_layer(params, prefix, x, layout: WindowLayout, shift: int, cfg: PathConfig):
    """models/sst/sst_basic_block.py:26-61 (WindowAttention) + :85-102 (EncoderLayer, post-norm)."""
    feat = layout.to_windows(shift, x)
    pos = layout.to_windows(shift, layout.shifts[shift]["pos"])
    pad = layout.key_padding(shift)
    outs = {}
    for dl, f3 in feat.items():
        f = f3.permute(1, 0, 2)
        qk = f + pos[dl].permute(1, 0, 2)
        o, _ = F.multi_head_attention_forward(
            qk, qk, f, cfg.d_model, cfg.nhead,
            params[prefix + "win_attn.self_attn.in_proj_weight"],
            params[prefix + "win_attn.self_attn.in_proj_bias"],
            None, None, False, 0.0,
            params[prefix + "win_attn.self_attn.out_proj.weight"],
            params[prefix + "win_attn.self_attn.out_proj.bias"],
            training=True, key_padding_mask=pad[dl], need_weights=False)
        outs[dl] = o.permute(1, 0, 2)
    src2 = layout.to_flat(shift, outs)
    d = cfg.d_model
    x = F.layer_norm(x + src2, (d,), params[prefix + "norm1.weight"], params[prefix + "norm1.bias"])
    h = F.gelu(F.linear(x, params[prefix + "linear1.weight"], params[prefix + "linear1.bias"]))
    src2 = F.linear(h, params[prefix + "linear2.weight"], params[prefix + "linear2.bias"])
    return F.layer_norm(x + src2, (d,), params[prefix + "norm2.weight"], params[prefix + "norm2.bias"])


def shift_block(params, prefix, x, layout, cfg):
    """models/sst/sst_basic_block.py:119-147 — layer 0 on shift-0 windows, layer 1 on shift-1."""
    for j in range(2):
        x = sra_layer(params, f"{prefix}encoder_list.{j}.", x, layout, j % len(cfg.shifts), cfg)
    return x


def backbone_forward(params, vis_feat, vis_coors, mask_coors, cfg: PathConfig,
                     prefix="backbone.", trace=None):
    """backbones/…top_only.py:136-141,199-303."""
    enc_layout = WindowLayout(vis_coors, cfg)
    x = vis_feat
    for i in range(cfg.enc_blocks):
        x = shift_block(params, f"{prefix}encoder_blocks.{i}.", x, enc_layout, cfg)
        if trace is not None:
            trace[f"enc{i}"] = x
    n_vis = vis_coors.shape[0]
    tokens = torch.cat([x, params[prefix + "mask_token"].repeat(mask_coors.shape[0], 1)], dim=0)
    dec_layout = WindowLayout(np.concatenate([vis_coors, mask_coors], axis=0), cfg)
    cen, den = tokens, tokens
    for i in range(cfg.dec_blocks):
        cen = shift_block(params, f"{prefix}decoder_centroid_blocks.{i}.", cen, dec_layout, cfg)
    for i in range(cfg.dec_blocks):
        den = shift_block(params, f"{prefix}decoder_density_blocks.{i}.", den, dec_layout, cfg)
    cen, den = cen[n_vis:], den[n_vis:]

    def head(name, t):
        return F.linear(t, params[f"{prefix}{name}.weight"], params[f"{prefix}{name}.bias"])
    return dict(reg_low=head("decoder_pred_low", cen).view(-1, cfg.slots_low, 3),
                reg_med=head("decoder_pred_med", cen).view(-1, cfg.slots_med, 3),
                reg_top=head("decoder_pred_top", cen),
                nor_top=head("decoder_pred_density_top", den),
                cls_low=head("cls_pred_low", cen).view(-1, cfg.slots_low, 2),
                cls_med=head("cls_pred_med", cen).view(-1, cfg.slots_med, 2),
                enc_layout=enc_layout, dec_layout=dec_layout)


# ---------------------------------------------------------------------------
# a22  losses
# ---------------------------------------------------------------------------
def masked_mse(pred, target, weight):
    """detectors/…_ssl.py:853-870 — mean over xyz, then sum / rows, times weight."""
    per_row = ((pred - target) ** 2).mean(dim=-1)
    return per_row.sum() / per_row.shape[0] * weight


def occupancy_bce(logits, occupied):
    """mmdet 2.20.0 CrossEntropyLoss(use_sigmoid=True) restated (call at …_ssl.py:894-895):
    labels one-hot expanded to the 2 channels, BCE-with-logits, mean over all elements."""
    onehot = F.one_hot(occupied.long(), 2).to(logits.dtype)
    return F.binary_cross_entropy_with_logits(logits, onehot, reduction="mean")


def losses(pred, tgt, cfg: PathConfig):
    """detectors/…_ssl.py:837-902 with mse_loss=True, cls_sub_voxel=True, nor_usr_sml1=None."""
    lm = torch.from_numpy(tgt["tgt_low_mask"]).reshape(-1)
    mm = torch.from_numpy(tgt["tgt_med_mask"]).reshape(-1)
    t_low = torch.from_numpy(tgt["tgt_low"]).reshape(-1, 3)[lm]
    t_med = torch.from_numpy(tgt["tgt_med"]).reshape(-1, 3)[mm]
    return dict(
        loss_curv_around=masked_mse(pred["nor_top"], torch.from_numpy(tgt["tgt_normal"]), cfg.loss_nor),
        loss_centroid_low=masked_mse(pred["reg_low"].reshape(-1, 3)[lm], t_low, cfg.loss_low),
        loss_centroid_med=masked_mse(pred["reg_med"].reshape(-1, 3)[mm], t_med, cfg.loss_med),
        loss_centroid_top=masked_mse(pred["reg_top"], torch.from_numpy(tgt["tgt_top"]), cfg.loss_top),
        loss_cls_low=occupancy_bce(pred["cls_low"].reshape(-1, 2), lm) * cfg.cls_low,
        loss_cls_med=occupancy_bce(pred["cls_med"].reshape(-1, 2), mm) * cfg.cls_med)


def forward_train(params, frames, cfg: PathConfig, ids_keep, ids_mask, trace=None, normal_override=None):
    """detectors/…_ssl.py:126-166 end to end.  ``params``: reference state_dict names -> tensors.
    ``normal_override`` ([V,3]) replaces the per-pillar normals before the loss — the parity harness
    uses it to sign-align / substitute the SVD-backend-dependent normals."""
    tgt = geometric_targets(frames, cfg, ids_mask)
    if normal_override is not None:
        tgt["tgt_normal"] = np.asarray(normal_override, np.float32)[np.asarray(ids_mask)]
    pts = torch.from_numpy(np.concatenate(frames, axis=0).astype(np.float32))
    feats, rows, inv = vfe_forward(params, pts, tgt["coors_top"], cfg)
    if trace is not None:
        trace["voxel_features"] = feats
    pred = backbone_forward(params, feats[torch.from_numpy(np.asarray(ids_keep))],
                            rows[np.asarray(ids_keep)], rows[np.asarray(ids_mask)], cfg, trace=trace)
    return losses(pred, tgt, cfg), pred, tgt


def init_params(cfg: PathConfig, seed=0):
    """Random-init parameters with the reference's names/shapes:
    xavier_uniform on backbone matrices (backbones/…top_only.py:318-321), zeros mask token."""
    g = torch.Generator().manual_seed(seed)
    d, f = cfg.d_model, cfg.ffn
    p = {}

    def xavier(*shape):
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    def small(*shape):
        return (torch.rand(*shape, generator=g) * 2 - 1) * 0.05
    groups = ([f"encoder_blocks.{i}" for i in range(cfg.enc_blocks)] +
              [f"decoder_centroid_blocks.{i}" for i in range(cfg.dec_blocks)] +
              [f"decoder_density_blocks.{i}" for i in range(cfg.dec_blocks)])
    for grp in groups:
        for j in range(2):
            k = f"backbone.{grp}.encoder_list.{j}."
            p[k + "win_attn.self_attn.in_proj_weight"] = xavier(3 * d, d)
            p[k + "win_attn.self_attn.in_proj_bias"] = small(3 * d)
            p[k + "win_attn.self_attn.out_proj.weight"] = xavier(d, d)
            p[k + "win_attn.self_attn.out_proj.bias"] = small(d)
            p[k + "linear1.weight"], p[k + "linear1.bias"] = xavier(f, d), small(f)
            p[k + "linear2.weight"], p[k + "linear2.bias"] = xavier(d, f), small(d)
            for nrm in ("norm1", "norm2"):
                p[k + nrm + ".weight"] = 1 + small(d)
                p[k + nrm + ".bias"] = small(d)
    p["backbone.mask_token"] = small(1, d)
    for name, n_out in (("decoder_pred_low", cfg.slots_low * 3), ("decoder_pred_med", cfg.slots_med * 3),
                        ("decoder_pred_top", 3), ("decoder_pred_density_top", 3),
                        ("cls_pred_low", cfg.slots_low * 2), ("cls_pred_med", cfg.slots_med * 2)):
        p[f"backbone.{name}.weight"], p[f"backbone.{name}.bias"] = xavier(n_out, d), small(n_out)
    p["voxel_encoder.vfe_layers.0.linear.weight"] = xavier(64, 11)
    p["voxel_encoder.vfe_layers.1.linear.weight"] = xavier(128, 128)
    for i, c in ((0, 64), (1, 128)):
        p[f"voxel_encoder.vfe_layers.{i}.norm.weight"] = 1 + small(c)
        p[f"voxel_encoder.vfe_layers.{i}.norm.bias"] = small(c)
    return p


# ---------------------------------------------------------------------------
# N2  data step in front of the path: augmentation + range filter
# ---------------------------------------------------------------------------
def augment_filter(frames, params, pc_range):
    """GlobalRotScaleTrans (translation std 0) -> RandomFlip3D -> PointsRangeFilter on each frame
    (datasets/pipelines/transforms_3d.py:670-718,95-123,849-883 via core/points/base_points.py:139-179,263-269,207-229
    and lidar_points.py:28-33).  params: [B,4] float32 rows (cos, sin, scale, flip bits: 1 horizontal, 2 vertical).
    fp32 arithmetic with every product and sum rounded separately:
        x' = (x c - y s) * scale,  y' = (x s + y c) * scale,  z' = z * scale;  strict  min < p' < max.
    Returns the list of filtered frames (input order kept, further channels untouched)."""
    lo = np.asarray(pc_range[:3], np.float32)
    hi = np.asarray(pc_range[3:], np.float32)
    out = []
    for f, (c, s, sc, fl) in zip(frames, np.asarray(params, np.float32)):
        f = np.asarray(f, np.float32)
        x, y, z = f[:, 0], f[:, 1], f[:, 2]
        xr = (x * c - y * s) * sc
        yr = (x * s + y * c) * sc
        zr = z * sc
        if int(fl) & 1:
            yr = -yr
        if int(fl) & 2:
            xr = -xr
        keep = (xr > lo[0]) & (yr > lo[1]) & (zr > lo[2]) & (xr < hi[0]) & (yr < hi[1]) & (zr < hi[2])
        g = f.copy()
        g[:, 0], g[:, 1], g[:, 2] = xr, yr, zr
        out.append(np.ascontiguousarray(g[keep]))
    return out


# ---------------------------------------------------------------------------
# N1  fine-tune consumer: SSTInputLayer (voxel drop) + SSTSecondPretrainedv1
# ---------------------------------------------------------------------------
def inner_win_inds_stable(win: np.ndarray):
    """get_inner_win_inds (middle_encoders/sst_input_layer.py:135-171) with a STABLE sort: a voxel's rank inside its
    window is its order of appearance.  The reference calls torch.sort without stable=True, so which voxel gets
    which rank is implementation-defined there (its docstring says so); every choice is a valid instance."""
    order = np.argsort(win, kind="stable")
    srt = win[order]
    inner = np.empty(win.size, np.int64)
    inner[order] = np.arange(win.size) - np.searchsorted(srt, srt, side="left")
    return inner


def input_layer_drop(coors: np.ndarray, cfg: PathConfig, inner_fn=inner_win_inds_stable):
    """middle_encoders/sst_input_layer.py:51-103 with shuffle_voxels=False: region grouping (:335-366), drop per
    shift (:211-236; bucket rule lower < n <= upper, :222; the voxels ranked below the budget stay), shift 1 on the
    survivors of shift 0 (:252-262).  ``inner_fn`` supplies the in-window rank (make_golden_n1 plugs in the
    reference's own get_inner_win_inds to pin everything else bit-exactly).
    -> (voxel_keep_inds, [level_shift0, level_shift1] of the kept voxels)."""
    def single(win):
        cnt = np.bincount(win)[win]
        inner = inner_fn(win)
        target = np.zeros(win.size, np.int64)
        lvl = np.full(win.size, -1, np.int64)
        for dl, info in cfg.drop_info.items():
            lo, hi = info["drop_range"]
            m = (cnt > lo) & (cnt <= hi)
            target[m], lvl[m] = info["max_tokens"], dl
        return inner < target, lvl

    keep = np.arange(coors.shape[0])
    keep0, lvl0 = single(window_partition(coors, cfg, 0)[0])
    keep, lvl0 = keep[keep0], lvl0[keep0]
    if len(cfg.shifts) == 1:
        return keep, [lvl0]
    keep1, lvl1 = single(window_partition(coors, cfg, 1)[0][keep0])
    return keep[keep1], [lvl0[keep1], lvl1[keep1]]


def recover_bev(feat: torch.Tensor, coors: np.ndarray, batch_size: int, ny: int, nx: int):
    """backbones/sst_second_pretrained_v1.py:246-276 -> [B, C, ny, nx]."""
    c = feat.shape[1]
    canvas = feat.new_zeros((batch_size, c, ny * nx))
    b = torch.from_numpy(coors[:, 0].astype(np.int64))
    at = torch.from_numpy((coors[:, 2].astype(np.int64) * nx + coors[:, 3]))
    canvas = _bev_put(canvas, b, at, feat)
    return canvas.view(batch_size, c, ny, nx)


def _bev_put(canvas, b, at, feat):
    flat = canvas.permute(0, 2, 1).reshape(-1, canvas.shape[1])
    flat = flat.index_put((b * canvas.shape[2] + at,), feat)
    return flat.view(canvas.shape[0], canvas.shape[2], canvas.shape[1]).permute(0, 2, 1).contiguous()


def second_stages(params, x, layer_nums, strides, eps, prefix="backbone.conv_blocks."):
    """backbones/sst_second_pretrained_v1.py:137-166,208-213: per stage a strided 3x3 conv + BN + ReLU followed by
    layer_num x (3x3 conv + BN + ReLU); BN in training mode (batch statistics), no conv bias."""
    outs = []
    for i, (n, stride) in enumerate(zip(layer_nums, strides)):
        for j in range(n + 1):
            w = params[f"{prefix}{i}.{3 * j}.weight"]
            x = F.conv2d(x, w, None, stride=stride if j == 0 else 1, padding=1)
            x = F.batch_norm(x, None, None, params[f"{prefix}{i}.{3 * j + 1}.weight"],
                             params[f"{prefix}{i}.{3 * j + 1}.bias"], training=True, eps=eps)
            x = F.relu(x)
        outs.append(x)
    return outs


def sst_second_forward(params, voxel_feat, coors, batch_size, cfg: PathConfig, n_blocks, output_shape, layer_nums,
                       strides, bn_eps=1e-3):
    """SSTInputLayer.forward + SSTSecondPretrainedv1.forward (:170-214) on pillar rows ``coors`` (b,z,y,x)."""
    keep, levels = input_layer_drop(coors, cfg)
    x, kept = voxel_feat[torch.from_numpy(keep)], coors[keep]
    layout = WindowLayout(kept, cfg, levels=levels)
    for i in range(n_blocks):
        x = shift_block(params, f"backbone.encoder_blocks.{i}.", x, layout, cfg)
    canvas = recover_bev(x, kept, batch_size, *output_shape)
    return keep, levels, layout, x, second_stages(params, canvas, layer_nums, strides, bn_eps)


def init_params_second(cfg: PathConfig, n_blocks, conv_in, conv_out, layer_nums, seed=0):
    """Random parameters with SSTSecondPretrainedv1's names and shapes."""
    g = torch.Generator().manual_seed(seed)
    d, f = cfg.d_model, cfg.ffn
    p = {}

    def uni(bound, *shape):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound
    for i in range(n_blocks):
        for j in range(2):
            k = f"backbone.encoder_blocks.{i}.encoder_list.{j}."
            p[k + "win_attn.self_attn.in_proj_weight"] = uni(math.sqrt(6.0 / (4 * d)), 3 * d, d)
            p[k + "win_attn.self_attn.in_proj_bias"] = uni(0.05, 3 * d)
            p[k + "win_attn.self_attn.out_proj.weight"] = uni(math.sqrt(3.0 / d), d, d)
            p[k + "win_attn.self_attn.out_proj.bias"] = uni(0.05, d)
            p[k + "linear1.weight"], p[k + "linear1.bias"] = uni(math.sqrt(6.0 / (d + f)), f, d), uni(0.05, f)
            p[k + "linear2.weight"], p[k + "linear2.bias"] = uni(math.sqrt(6.0 / (d + f)), d, f), uni(0.05, d)
            for nrm in ("norm1", "norm2"):
                p[k + nrm + ".weight"], p[k + nrm + ".bias"] = 1 + uni(0.05, d), uni(0.05, d)
    cin = [conv_in, *conv_out[:-1]]
    for i, n in enumerate(layer_nums):
        for j in range(n + 1):
            ci = cin[i] if j == 0 else conv_out[i]
            p[f"backbone.conv_blocks.{i}.{3 * j}.weight"] = uni(math.sqrt(2.0 / (ci * 9)), conv_out[i], ci, 3, 3)
            p[f"backbone.conv_blocks.{i}.{3 * j + 1}.weight"] = 1 + uni(0.05, conv_out[i])
            p[f"backbone.conv_blocks.{i}.{3 * j + 1}.bias"] = uni(0.05, conv_out[i])
    return p
