"""SSTInputLayer — mirror of mmdet3d/models/middle_encoders/sst_input_layer.py:14-103 (same registry key, constructor
arguments, ``forward(voxel_feat, coors, batch_size)`` and the same three outputs), the entry of the fine-tune consumer
(SURVEY.md §8(f) N1).

What the reference does with a randperm, two sorts, bincounts and uniques per shift is here two launches of
``geomae_window_drop`` over the occupancy bitmap (csrc/window_csr.cu): count each window, pick its bucket with
``lower < n <= upper`` (:222), keep ``max_tokens`` of its voxels, shift 1 on the survivors of shift 0 (:252-262).
The survivors' CSR window layout (``voxel_info['window_layout']``) is what this package's SRA kernels consume; the
reference-format outputs (``flat2win_inds_list``, ``batch_win_inds_shift*``, ``coors_in_win_shift*``,
``voxel_drop_level_shift*``, ``voxel_keep_inds``) are derived from it for callers written against the reference.

Differences a caller can observe, both inside what the reference itself leaves unspecified:
  * ``shuffle_voxels=True``: rows are NOT physically permuted; the drop picks a uniformly random subset per window
    (hash of a seed drawn from torch's CPU generator), which is the only effect the reference's shuffle has (:66-74).
  * the in-window slot of a voxel (``flat2window_inds % max_tokens``) follows cell order, not sort order — the
    reference documents its own order as unstable (:135)."""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import lib as L
from .registry import MIDDLE_ENCODERS
from .voxel import VoxelGeometry
from .windows import WindowLayout, WindowSpec, coors_bitmap


@MIDDLE_ENCODERS.register_module()
class SSTInputLayer(nn.Module):
    def __init__(self, drop_info, shifts_list, window_shape, point_cloud_range, voxel_size, shuffle_voxels=True,
                 debug=True):
        super().__init__()
        self.fp16_enabled = False
        self.meta_drop_info = drop_info
        self.shifts_list = [tuple(s) for s in shifts_list]
        self.point_cloud_range, self.voxel_size = point_cloud_range, tuple(voxel_size)
        self.shuffle_voxels, self.debug = shuffle_voxels, debug
        self.window_shape = tuple(window_shape)
        for sx, sy in self.shifts_list:       # :350
            assert sx in (0, self.window_shape[0] // 2) and sy in (0, self.window_shape[1] // 2), \
                "shift must be 0 or half a window"
        self.spec = WindowSpec(self.window_shape, self.shifts_list)
        self.geom = VoxelGeometry(tuple(point_cloud_range), self.voxel_size, self.voxel_size, self.voxel_size,
                                  (1, 1, 1), (1, 1, 1))

    def set_drop_info(self):
        """:378-389 — (training, test) tuple resolved once, on first use."""
        if hasattr(self, "drop_info"):
            return
        meta = self.meta_drop_info
        self.drop_info = (meta[0] if self.training else meta[1]) if isinstance(meta, tuple) else meta
        levels = sorted(self.drop_info)
        assert levels == list(range(len(levels))) and len(levels) <= 8, "drop levels must be 0..n-1, n <= 8"
        arr = lambda f: (C.c_int32 * len(levels))(*[f(self.drop_info[l]) for l in levels])   # noqa: E731
        self._levels = (len(levels), arr(lambda d: d["max_tokens"]), arr(lambda d: d["drop_range"][0]),
                        arr(lambda d: d["drop_range"][1]))

    @torch.no_grad()
    def drop(self, coors, batch_size):
        """-> (keep [n] uint8, level [n_shifts, n] int32) for token rows ``coors`` (int32 [n,4], device)."""
        n = coors.shape[0]
        io, tok_of_pillar, alive = coors_bitmap(self.geom, coors, batch_size)
        keep = torch.empty(max(n, 1), dtype=torch.uint8, device=coors.device)
        level = torch.full((self.spec.n_shifts, max(n, 1)), -1, dtype=torch.int32, device=coors.device)
        seed = int(torch.randint(1, 2 ** 62, (1,)).item()) if self.shuffle_voxels else 0
        nl, mx, lo, up = self._levels
        L.run("window_drop", C.byref(self.geom.cstruct), C.byref(self.spec.cstruct), C.byref(io), L.ptr(tok_of_pillar),
              n, nl, mx, lo, up, seed, L.ptr(keep), L.ptr(level), L.stream_ptr(coors.device))
        del alive
        return keep[:n], level[:, :n]

    @torch.no_grad()
    def get_flat2win_inds(self, layout: WindowLayout, shift: int, voxel_drop_lvl):
        """:105-133 — {level: (window_slot * max_tokens + in_window_slot, where(level mask))} from the CSR arrays."""
        tok_win = layout.tok_win[shift].long()
        inner = layout.tok_pos[shift].long() - layout.win_ptr[shift].long()[tok_win]
        win_level = torch.full((layout.max_windows + 1,), -1, dtype=torch.long, device=tok_win.device)
        win_level[tok_win] = voxel_drop_lvl
        out = {}
        for dl, info in self.drop_info.items():
            mask = voxel_drop_lvl == dl
            if not mask.any():
                continue
            slot_of_win = torch.cumsum((win_level == dl).long(), 0) - 1
            out[dl] = (slot_of_win[tok_win[mask]] * info["max_tokens"] + inner[mask], torch.where(mask))
        return out

    def forward(self, voxel_feat, coors, batch_size):
        """voxel_feat [N, C], coors [N, 4] (b, z, y, x) -> (voxel_feat of the kept voxels, flat2win_inds_list,
        voxel_info) — :51-103."""
        self.set_drop_info()
        L.require_cuda(voxel_feat, "voxel_feat")
        coors32 = coors.to(torch.int32).contiguous()
        keep, level = self.drop(coors32, batch_size)
        keep_inds = torch.nonzero(keep).squeeze(1)      # output length: the one host sync, as in the reference (:82-85)
        dropped = keep_inds.shape[0] != coors32.shape[0]
        if dropped:
            voxel_feat, coors32 = voxel_feat.index_select(0, keep_inds), coors32.index_select(0, keep_inds)
            level = level.index_select(1, keep_inds)
        layout = WindowLayout.from_coors(self.spec, self.geom, coors32, batch_size)
        wy = self.window_shape[1]
        voxel_info = dict(coors=coors32.long(), voxel_keep_inds=keep_inds, window_layout=layout)
        flat2win_inds_list = []
        for i in range(self.spec.n_shifts):
            cell = layout.tok_cell[i].long()
            voxel_info[f"batch_win_inds_shift{i}"] = layout.win_id[i].long()[layout.tok_win[i].long()]
            voxel_info[f"coors_in_win_shift{i}"] = torch.stack([cell // wy, cell % wy], dim=-1)
            voxel_info[f"voxel_drop_level_shift{i}"] = level[i].long()
            flat2win_inds_list.append(self.get_flat2win_inds(layout, i, voxel_info[f"voxel_drop_level_shift{i}"]))
        return voxel_feat, flat2win_inds_list, voxel_info
