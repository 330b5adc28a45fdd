"""GPU-side data step in front of the voxel scatter.

The reference augments and range-filters every sample on the data-loader workers' CPU cores
(``GlobalRotScaleTrans``, ``RandomFlip3D``, ``PointsRangeFilter``: datasets/pipelines/transforms_3d.py:95-123,
670-768,849-883).  Here the raw frames of a batch go to the device once and ONE C-ABI call transforms, filters and
compacts all of them (``geomae_augment_filter``); the random draws stay on the host and follow the reference's order.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import lib as L


@dataclass
class Augmentation:
    """One sample's draws: rotation angle about z (rad), isotropic scale, BEV flips."""
    rotation: float = 0.0
    scale: float = 1.0
    flip_horizontal: bool = False   # y -> -y  (lidar_points.py:30-31)
    flip_vertical: bool = False     # x -> -x  (lidar_points.py:32-33)


def draw_augmentation(rng: np.random.RandomState, rot_range=(-0.3925, 0.3925), scale_ratio_range=(0.95, 1.05),
                      flip_ratio_bev_horizontal=0.5, flip_ratio_bev_vertical=0.5) -> Augmentation:
    """The reference's draws in the reference's order: rotation (transforms_3d.py:679-680), scale (:728-730),
    translation noise (:660, std 0 in the GeoMAE configs: drawn and discarded), then the two flip decisions
    (RandomFlip3D.__call__, :143-152)."""
    rot = rng.uniform(rot_range[0], rot_range[1])
    scale = rng.uniform(scale_ratio_range[0], scale_ratio_range[1])
    rng.normal(scale=0.0, size=3)
    flip_h = bool(rng.rand() < flip_ratio_bev_horizontal)
    flip_v = bool(rng.rand() < flip_ratio_bev_vertical)
    return Augmentation(float(rot), float(scale), flip_h, flip_v)


def frame_params(augs) -> torch.Tensor:
    """[n_frames, 4] float32 host tensor: cos, sin, scale, flip bits.  sin / cos are evaluated in float32 by torch, as
    ``BasePoints.rotate`` does (base_points.py:147-157); the scale is rounded to float32 like the in-place ``*=``."""
    rot = torch.tensor([a.rotation for a in augs], dtype=torch.float32)
    out = torch.empty((len(augs), 4), dtype=torch.float32)
    out[:, 0] = torch.cos(rot)
    out[:, 1] = torch.sin(rot)
    out[:, 2] = torch.tensor([a.scale for a in augs], dtype=torch.float32)
    out[:, 3] = torch.tensor([int(a.flip_horizontal) + 2 * int(a.flip_vertical) for a in augs], dtype=torch.float32)
    return out


def augment_filter(points: torch.Tensor, frame_offsets: torch.Tensor, augs, point_cloud_range):
    """points [N, C>=3] float32 CUDA (frames concatenated), frame_offsets [B+1] int32 CUDA ->
    (filtered points [N, C] of which the first ``offsets[-1]`` rows are valid, new offsets [B+1] int32 CUDA).
    No host synchronisation: downstream kernels read the device-side offsets."""
    L.require_cuda(points, "points")
    L.require_cuda(frame_offsets, "frame_offsets")
    if points.dtype != torch.float32 or frame_offsets.dtype != torch.int32:
        raise RuntimeError("augment_filter: points must be float32 and frame_offsets int32")
    points = points.contiguous()
    n, stride = points.shape
    n_frames = frame_offsets.numel() - 1
    if len(augs) != n_frames:
        raise RuntimeError(f"augment_filter: {len(augs)} augmentations for {n_frames} frames")
    dev = points.device
    params = frame_params(augs).to(dev, non_blocking=True)
    out = torch.empty_like(points)
    out_off = torch.empty(n_frames + 1, dtype=torch.int32, device=dev)
    n_tmp = (n + 1023) // 1024 + 1
    tmp = torch.empty(n_tmp, dtype=torch.int32, device=dev)
    rng = point_cloud_range
    L.run("augment_filter", L.ptr(points), n, stride, L.ptr(frame_offsets), n_frames, L.ptr(params),
          L.f3(rng[:3]), L.f3(rng[3:]), L.ptr(out), L.ptr(out_off), L.ptr(tmp), C.c_int64(n_tmp), L.stream_ptr(dev))
    return out, out_off
