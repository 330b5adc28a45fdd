"""CSR window layouts (both shifts) for a token set — host side of csrc/window_csr.cu.

Replaces the per-forward bookkeeping of ``MultiMAESSTSPChoose.get_voxel_info``
(backbones/multi_mae_sst_spearate_top_only.py:143-196): window_partition, drop levels,
flat2win indices, key-padding masks and padded position embeddings.  Nothing here
synchronises with the host; window counts stay on the device.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L
from .voxel import PillarBatch, VoxelGeometry


def coors_bitmap(geom: VoxelGeometry, coors: torch.Tensor, batch_size: int):
    """Occupancy bitmap + ranks of explicit token rows ``coors [n,4] = (b,z,y,x)``: (ScatterIO view, tok_of_pillar,
    buffers to keep alive) — what the window kernels need when no scatter result is at hand."""
    L.require_cuda(coors, "coors")
    dev = coors.device
    coors = coors.to(torch.int32).contiguous()
    n = coors.shape[0]
    gx, gy, _ = geom.grid
    n_words = (batch_size * gx * gy + 31) // 32
    i32 = dict(dtype=torch.int32, device=dev)
    bitmap, word_rank = torch.empty(n_words, **i32), torch.empty(n_words, **i32)
    scan_tmp, counts = torch.empty(3 * 16384, **i32), torch.empty(4, **i32)     # counts: cleared by the C call
    tok_of_pillar = torch.empty(max(n, 1), **i32)
    L.run("coors_bitmap", C.byref(geom.cstruct), L.ptr(coors), n, batch_size, L.ptr(bitmap), L.ptr(word_rank),
          L.ptr(scan_tmp), L.ptr(counts), L.ptr(tok_of_pillar), L.stream_ptr(dev))
    io = L.ScatterIO()
    io.n_frames, io.bitmap, io.word_rank = batch_size, L.ptr(bitmap), L.ptr(word_rank)
    return io, tok_of_pillar, (bitmap, word_rank, scan_tmp, counts, tok_of_pillar, coors)


class WindowSpec:
    """window_shape + shifts_list of the backbone config (…6x_1e-5.py:15,57)."""

    def __init__(self, window_shape, shifts_list):
        assert 1 <= len(shifts_list) <= 2
        self.window_shape = tuple(window_shape)
        self.shifts_list = [tuple(s) for s in shifts_list]
        sx = [s[0] for s in self.shifts_list] + [0] * (2 - len(shifts_list))
        sy = [s[1] for s in self.shifts_list] + [0] * (2 - len(shifts_list))
        self.cstruct = L.WindowCfg(window_shape[0], window_shape[1], len(shifts_list),
                                   (C.c_int32 * 2)(*sx), (C.c_int32 * 2)(*sy))
        self.n_shifts = len(shifts_list)

    def candidates(self, geom: VoxelGeometry, n_frames: int) -> int:
        n = C.c_int32()
        L.run("window_candidates", C.byref(geom.cstruct), C.byref(self.cstruct), n_frames,
                                                 C.byref(n), None, None)
        return n.value


class WindowLayout:
    """Device-side CSR windows of one token set.  ``shift(i)`` returns the arrays of shift i:
    n_windows [1] i32 (device), win_ptr, win_tok, tok_cell, tok_win, tok_pos, max_windows (host bound)."""

    def __init__(self, spec: WindowSpec, geom: VoxelGeometry, n_frames: int, n_tokens: int, device):
        self.spec, self.geom, self.n_tokens, self.n_frames = spec, geom, n_tokens, n_frames
        self.n_cand = spec.candidates(geom, n_frames)
        ns = spec.n_shifts
        i32 = dict(dtype=torch.int32, device=device)
        self.ptr_stride = self.n_cand + 1
        nt = max(n_tokens, 1)
        self._scratch = torch.empty((3, ns, self.n_cand), **i32)
        self.n_windows = torch.empty(ns, **i32)        # written by k_win_scan for every shift
        self.win_ptr = torch.empty((ns, self.ptr_stride), **i32)
        self.win_id = torch.empty((ns, self.ptr_stride), **i32)
        self.win_tok = torch.empty((ns, nt), **i32)
        self.tok_cell = torch.empty((ns, nt), **i32)
        self.tok_win = torch.empty((ns, nt), **i32)
        self.tok_pos = torch.empty((ns, nt), **i32)
        self.io = L.WindowIO(self.ptr_stride, L.ptr(self._scratch[0]), L.ptr(self._scratch[1]),
                             L.ptr(self._scratch[2]), L.ptr(self.n_windows), L.ptr(self.win_ptr),
                             L.ptr(self.win_id), L.ptr(self.win_tok), L.ptr(self.tok_cell), L.ptr(self.tok_win),
                             L.ptr(self.tok_pos))
        self.max_windows = min(self.n_cand, nt)

    @classmethod
    def from_pillars(cls, spec, pb: PillarBatch, rows: torch.Tensor):
        """Token i = pillar ``rows[i]`` of an existing scatter result (bitmap re-used)."""
        dev = rows.device
        self = cls(spec, pb.geom, pb.n_frames, rows.shape[0], dev)
        rows = rows.to(torch.int64).contiguous()
        tok_of_pillar = torch.empty(max(pb.n_pillars, 1), dtype=torch.int32, device=dev)
        s = L.stream_ptr(dev)
        L.run("token_map", L.ptr(rows), rows.shape[0], L.ptr(tok_of_pillar), pb.n_pillars, s)
        L.run("window_csr", C.byref(pb.geom.cstruct), C.byref(spec.cstruct), C.byref(pb.io),
                                          L.ptr(tok_of_pillar), rows.shape[0], C.byref(self.io), s)
        self._keep = (tok_of_pillar, rows)
        return self

    @classmethod
    def from_coors(cls, spec, geom: VoxelGeometry, coors: torch.Tensor, batch_size: int):
        """Token i sits at ``coors[i] = (b,z,y,x)`` (unique cells), the reference's call form."""
        io, tok_of_pillar, keep = coors_bitmap(geom, coors, batch_size)
        n = coors.shape[0]
        self = cls(spec, geom, batch_size, n, coors.device)
        L.run("window_csr", C.byref(geom.cstruct), C.byref(spec.cstruct), C.byref(io),
                                          L.ptr(tok_of_pillar), n, C.byref(self.io), L.stream_ptr(coors.device))
        self._keep = keep
        return self

    def hand_over(self, stream):
        """Built under another stream: from now on ``stream`` uses these buffers (see PillarBatch.hand_over)."""
        for t in (self._scratch, self.n_windows, self.win_ptr, self.win_id, self.win_tok, self.tok_cell, self.tok_win,
                  self.tok_pos) + tuple(t for t in getattr(self, "_keep", ()) if torch.is_tensor(t)):
            t.record_stream(stream)

    def shift(self, i: int):
        return dict(n_windows=self.n_windows[i:i + 1], win_ptr=self.win_ptr[i], win_tok=self.win_tok[i],
                    tok_cell=self.tok_cell[i], tok_win=self.tok_win[i], tok_pos=self.tok_pos[i],
                    max_windows=self.max_windows)


_POS_TABLES = {}


def pos_table(window_shape, d_model, temperature, device) -> torch.Tensor:
    """[win_x*win_y, d_model] sinusoid table (…top_only.py:361-394), computed once per device."""
    key = (tuple(window_shape), d_model, float(temperature), str(device))
    if key not in _POS_TABLES:
        t = torch.empty((window_shape[0] * window_shape[1], d_model), dtype=torch.float32, device=device)
        L.run("pos_table", window_shape[0], window_shape[1], d_model, float(temperature), L.ptr(t),
                                         L.stream_ptr(device))
        _POS_TABLES[key] = t
    return _POS_TABLES[key]
