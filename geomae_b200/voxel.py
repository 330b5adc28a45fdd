"""Host-side mirror of the reference's voxel ops, routed to the sm_100a kernels.

``Voxelization`` keeps the reference signature (mmdet3d/ops/voxel/voxelize.py:63-112,
dynamic mode only: the hard-voxel layers are built but never called on the SSL path,
SURVEY.md §2.2).  ``scatter_frames`` is the fused replacement for the
voxelize x3 + torch.unique x6 + torch_scatter + scatter_add_ chain of
``MultiSubVoxelDynamicVoxelNetSSL.extract_feat`` (…_ssl.py:169-219).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch
from torch import nn

from . import lib as L


class Voxelization(nn.Module):
    """Dynamic voxelisation: ``points [N, C>=3] -> coors [N, 3] int32 (z, y, x)``, out-of-range
    points clamped into the edge voxels (this fork's behaviour, voxelization_cuda.cu:35-57)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)

    def forward(self, points: torch.Tensor) -> torch.Tensor:
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        if not (self.max_num_points == -1 or max_voxels == -1):
            raise NotImplementedError("hard voxelisation is outside the GeoMAE pretraining path")
        L.require_cuda(points, "points")
        if points.dtype != torch.float32:
            raise RuntimeError("points must be float32")
        points = points.contiguous()
        coors = torch.empty((points.shape[0], 3), dtype=torch.int32, device=points.device)
        rng = self.point_cloud_range
        L.run("dynamic_voxelize", L.ptr(points), points.shape[0], points.shape[1],
                                                L.f3(self.voxel_size), L.f3(rng[:3]), L.f3(rng[3:]),
                                                L.ptr(coors), L.stream_ptr(points.device))
        return coors

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range="
                f"{self.point_cloud_range}, max_num_points={self.max_num_points}, max_voxels={self.max_voxels})")


def grid_size(voxel_size, pc_range):
    """[x, y, z] grid of one scale = ceil((max-min)/size) in fp32 (voxelization_cuda.cu:375-377)."""
    out = (C.c_int32 * 3)()
    L.run("grid_size", L.f3(pc_range[:3]), L.f3(pc_range[3:]), L.f3(voxel_size), out)
    return list(out)


@dataclass
class VoxelGeometry:
    """Three-scale voxel geometry of a config (…6x_1e-5.py:14-24)."""
    pc_range: tuple
    voxel_size: tuple
    voxel_size_med: tuple
    voxel_size_low: tuple
    ratio_med: tuple   # z, y, x
    ratio_low: tuple   # z, y, x

    def __post_init__(self):
        self.cstruct = L.VoxelCfg(L.f3(self.pc_range[:3]), L.f3(self.pc_range[3:]), L.f3(self.voxel_size),
                                  L.f3(self.voxel_size_med), L.f3(self.voxel_size_low),
                                  (C.c_int32 * 3)(*self.ratio_med), (C.c_int32 * 3)(*self.ratio_low))
        self.slots_med = self.ratio_med[0] * self.ratio_med[1] * self.ratio_med[2]
        self.slots_low = self.ratio_low[0] * self.ratio_low[1] * self.ratio_low[2]
        self._grid = None

    @property
    def grid(self):
        if self._grid is None:
            self._grid = grid_size(self.voxel_size, self.pc_range)
        return self._grid


class PillarBatch:
    """Device-resident result of the fused voxelise+scatter stage for one batch of frames."""

    def __init__(self, geom: VoxelGeometry, points: torch.Tensor, frame_offsets: torch.Tensor, n_frames: int,
                 want_coors: bool = False):
        dev = points.device
        n = points.shape[0]
        self.geom, self.points, self.frame_offsets, self.n_frames = geom, points, frame_offsets, n_frames
        gx, gy, _ = geom.grid
        n_words = (n_frames * gx * gy + 31) // 32
        cap = max(n, 1)
        i32 = dict(dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.bitmap = torch.empty(n_words, **i32)
        self.word_rank = torch.empty(n_words, **i32)
        self.scan_tmp = torch.empty(3 * 16384, **i32)
        self.counts = torch.empty(4 + n_frames + 1, **i32)       # cleared by geomae_voxel_scatter
        self.pillar_coors = torch.empty((cap, 4), **i32)
        self.pillar_mean = torch.empty((cap, 4), **f32)
        self.point_pillar = torch.empty(cap, **i32)
        self.med_mask = torch.empty(cap, **i32)
        self.low_mask = torch.empty((cap, 4), **i32)
        self.med_ptr = torch.empty(cap + 1, **i32)
        self.low_ptr = torch.empty(cap + 1, **i32)
        self.med_mean = torch.empty((cap, 4), **f32)
        self.low_mean = torch.empty((cap, 4), **f32)
        self.coors_top = self.coors_med = self.coors_low = None
        if want_coors:
            self.coors_top = torch.empty((cap, 4), **i32)
            self.coors_med = torch.empty((cap, 4), **i32)
            self.coors_low = torch.empty((cap, 4), **i32)
        self.io = L.ScatterIO(
            L.ptr(points), L.ptr(frame_offsets), n, points.shape[1], n_frames, cap,
            L.ptr(self.bitmap), L.ptr(self.word_rank), L.ptr(self.scan_tmp), L.ptr(self.counts),
            L.ptr(self.pillar_coors), L.ptr(self.pillar_mean), L.ptr(self.point_pillar), L.ptr(self.med_mask),
            L.ptr(self.low_mask), L.ptr(self.med_ptr), L.ptr(self.low_ptr), L.ptr(self.med_mean),
            L.ptr(self.low_mean), L.ptr(self.coors_top), L.ptr(self.coors_med), L.ptr(self.coors_low))
        self._n = None

    def run(self, stream=None):
        """stream: raw CUDA stream handle to launch on (default: the caller's current / pinned stream)."""
        L.run("voxel_scatter", C.byref(self.geom.cstruct), C.byref(self.io),
              stream if stream is not None else L.stream_ptr(self.points.device))
        self._n = None
        return self

    def hand_over(self, stream):
        """The buffers were allocated (and filled) under another stream; tell the caching allocator that ``stream``
        uses them from now on, so that a later free is not recycled under work still queued there."""
        for t in (self.points, self.frame_offsets, self.bitmap, self.word_rank, self.scan_tmp, self.counts,
                  self.pillar_coors, self.pillar_mean, self.point_pillar, self.med_mask, self.low_mask, self.med_ptr,
                  self.low_ptr, self.med_mean, self.low_mean, self.coors_top, self.coors_med, self.coors_low):
            if t is not None:
                t.record_stream(stream)

    def sizes(self):
        """(n_pillars, n_med, n_low) — one device->host read of four ints."""
        if self._n is None:
            c = self.counts.tolist()
            if c[3]:
                raise RuntimeError("geomae_b200.voxel_scatter: pillar capacity exceeded")
            self._n = (c[0], c[1], c[2])
            self.frame_starts = c[4:]
        return self._n

    @property
    def n_pillars(self):
        return self.sizes()[0]

    def pillars_per_frame(self):
        """Host list of per-sample pillar counts (comes with the same device->host read as sizes())."""
        self.sizes()
        fs = self.frame_starts
        return [fs[b + 1] - fs[b] for b in range(self.n_frames)]

    def geom_targets(self, want_debug=False):
        """normal [V,3] f32 (z,y,x), curvature [V,3] f64 (+ cov6, singular, pair when want_debug)."""
        v = self.n_pillars
        dev = self.points.device
        normal = torch.empty((v, 3), dtype=torch.float32, device=dev)
        curv = torch.empty((v, 3), dtype=torch.float64, device=dev)
        cov6 = sing = pair = None
        if want_debug:
            cov6 = torch.empty((v, 6), dtype=torch.float32, device=dev)
            sing = torch.empty((v, 3), dtype=torch.float32, device=dev)
            pair = torch.empty((9, v), dtype=torch.int32, device=dev)
        L.run("geom_targets", C.byref(self.geom.cstruct), C.byref(self.io), v, L.ptr(normal),
                                            L.ptr(curv), L.ptr(cov6), L.ptr(sing), L.ptr(pair),
                                            L.stream_ptr(dev))
        return (normal, curv, cov6, sing, pair) if want_debug else (normal, curv)

    def dense_targets(self, rows: torch.Tensor, raw=False):
        """Dense slot targets of the selected pillar rows, reference layout:
        low [m,slots_low,3] + mask, med [m,slots_med,3] + mask, top [m,3] (all (z,y,x))."""
        dev = self.points.device
        rows = rows.to(torch.int64).contiguous()
        m = rows.shape[0]
        g = self.geom
        low = torch.empty((m, g.slots_low, 3), dtype=torch.float32, device=dev)
        low_m = torch.empty((m, g.slots_low), dtype=torch.uint8, device=dev)
        med = torch.empty((m, g.slots_med, 3), dtype=torch.float32, device=dev)
        med_m = torch.empty((m, g.slots_med), dtype=torch.uint8, device=dev)
        top = torch.empty((m, 3), dtype=torch.float32, device=dev)
        L.run("dense_targets", C.byref(g.cstruct), C.byref(self.io), L.ptr(rows), m, int(raw),
                                             L.ptr(low), L.ptr(low_m), L.ptr(med), L.ptr(med_m), L.ptr(top),
                                             L.stream_ptr(dev))
        return low, low_m.bool(), med, med_m.bool(), top


def scatter_frames(geom: VoxelGeometry, frames, want_coors=False, side=None) -> PillarBatch:
    """frames: list of [N_i, C] float32 CUDA tensors (or one concatenated tensor + offsets).

    ``side`` (a torch.cuda.Stream): concatenate, scatter and read the three totals back on that stream, then make the
    caller's stream wait for it.  The read then blocks the host for the scatter alone instead of for everything
    still queued on the compute stream (the previous step's backward and optimiser), which keeps the host enqueueing
    ahead of the device.  The caller guarantees that ``frames`` are complete or were produced on ``side``."""
    for f in frames:
        L.require_cuda(f, "points")
    offs = [0]
    for f in frames:
        offs.append(offs[-1] + f.shape[0])
    if side is None:
        points = torch.cat(frames, dim=0).contiguous()
        frame_offsets = torch.tensor(offs, dtype=torch.int32, device=points.device)
        return PillarBatch(geom, points, frame_offsets, len(frames), want_coors).run()
    dev = frames[0].device
    main = torch.cuda.current_stream(dev)
    with torch.cuda.stream(side):
        points = torch.cat(frames, dim=0).contiguous()
        frame_offsets = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
        pb = PillarBatch(geom, points, frame_offsets, len(frames), want_coors).run(side.cuda_stream)
        pb.sizes()
    main.wait_stream(side)
    pb.hand_over(main)
    return pb
