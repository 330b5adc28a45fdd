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
    translation noise (:660, std 0 in the GeoMAE configs: drawn and discarded), then RandomFlip3D.__call__: first the
    2-D flip draw of its mmdet base class (:138 -> mmdet 2.20 RandomFlip.__call__: one
    ``np.random.choice(['horizontal', None], p=[r, 1 - r])``, drawn and discarded — there are no images on this path),
    then the two BEV flip decisions (:143-152)."""
    rot = rng.uniform(rot_range[0], rot_range[1])
    scale = rng.uniform(scale_ratio_range[0], scale_ratio_range[1])
    rng.normal(scale=0.0, size=3)
    rng.choice(2, p=[flip_ratio_bev_horizontal, 1.0 - flip_ratio_bev_horizontal])
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


# ---------------------------------------------------------------------------------------------------------------------
# On-disk formats: nuScenes `.pcd.bin` sweeps and the `nuscenes_ssl_infos_*.pkl` index.
# Host-side numpy, statement for statement what the reference's loaders do (they are numpy too); the result is the raw
# [N, 5] float32 frame that goes to the device once and through `augment_filter`.
# ---------------------------------------------------------------------------------------------------------------------
def read_points_bin(path: str, load_dim: int = 5, use_dim=5) -> np.ndarray:
    """`LoadPointsFromFile` (datasets/pipelines/loading.py:382-427): a flat float32 file of `load_dim`-float records
    (nuScenes: x, y, z, intensity, ring index), columns `use_dim` kept."""
    if isinstance(use_dim, int):
        use_dim = list(range(use_dim))
    pts = np.fromfile(path, dtype=np.float32) if not path.endswith(".npy") else np.load(path)
    return pts.reshape(-1, load_dim)[:, use_dim]


def remove_close(points: np.ndarray, radius: float = 1.0) -> np.ndarray:
    """Points inside the square |x| < radius and |y| < radius around the sensor (ego-vehicle returns) are dropped
    (`LoadPointsFromMultiSweeps._remove_close`, loading.py:160-181)."""
    inside = (np.abs(points[:, 0]) < radius) & (np.abs(points[:, 1]) < radius)
    return points[~inside]


def _pick_sweeps(n_available: int, sweeps_num: int, test_mode: bool, rng) -> np.ndarray:
    """Which of the `n_available` earlier sweeps join the key frame: all of them when there are at most `sweeps_num`,
    the nearest `sweeps_num` at test time, a draw without replacement in training (loading.py:208-215)."""
    if n_available <= sweeps_num:
        return np.arange(n_available)
    if test_mode:
        return np.arange(sweeps_num)
    return rng.choice(n_available, sweeps_num, replace=False)


def _sweep_in_key_frame(raw: np.ndarray, sweep: dict, key_time: float, load_dim: int, drop_close: bool) -> np.ndarray:
    """One earlier sweep in the key frame's lidar coordinates, channel 4 = its time lag in seconds.  The arithmetic
    keeps the reference's types on purpose (loading.py:218-226): float32 points times the float64 sensor->lidar
    rotation, rounded back to float32, then the float64 translation added in place."""
    pts = np.array(raw, dtype=np.float32).reshape(-1, load_dim)
    if drop_close:
        pts = remove_close(pts)
    pts[:, :3] = pts[:, :3] @ sweep["sensor2lidar_rotation"].T
    pts[:, :3] += sweep["sensor2lidar_translation"]
    pts[:, 4] = key_time - sweep["timestamp"] / 1e6
    return pts


def load_multi_sweeps(points: np.ndarray, info: dict, sweeps_num: int = 9, load_dim: int = 5,
                      use_dim=(0, 1, 2, 3, 4), pad_empty_sweeps: bool = True, remove_close_points: bool = True,
                      test_mode: bool = False, rng=np.random, read=None) -> np.ndarray:
    """What `LoadPointsFromMultiSweeps` (loading.py:100-235) produces with the arguments of the GeoMAE pretraining
    config (…6x_1e-5.py:174-180): the key frame with its time channel zeroed, followed by up to `sweeps_num` earlier
    sweeps moved into the key frame's coordinates; a key frame without sweeps is padded with copies of itself.
    `info` = an entry of `NuScenesSSLIndex.get_data_info`; `rng` must offer `choice`; `read(path)` returns the flat
    float32 contents of a sweep file."""
    read = read or (lambda path: np.fromfile(path, dtype=np.float32))
    key = np.array(points, dtype=np.float32, copy=True)
    key[:, 4] = 0
    parts = [key]
    sweeps = info["sweeps"]
    if pad_empty_sweeps and not sweeps:
        filler = remove_close(key) if remove_close_points else key
        parts += [filler] * sweeps_num
    else:
        for i in _pick_sweeps(len(sweeps), sweeps_num, test_mode, rng):
            parts.append(_sweep_in_key_frame(read(sweeps[i]["data_path"]), sweeps[i], info["timestamp"], load_dim,
                                             remove_close_points))
    return np.concatenate(parts, axis=0)[:, list(use_dim)]


@dataclass
class RawSweeps:
    """One sample as it lies on disk: the raw [n_i, load_dim] float32 records of the key frame and of the chosen
    earlier sweeps (``arrays``) and what moves each of them into the key frame (``params`` [S, 16] float64, the
    layout of ``geomae_sweep_merge``).  ``FlatTrainer.train_step_from_host`` merges them on the device."""
    arrays: list
    params: np.ndarray

    @property
    def n_points(self):
        return sum(a.shape[0] for a in self.arrays)


def _segment_params(rotation=None, translation=None, time_lag=0.0, close_radius=-1.0):
    p = np.zeros(16, np.float64)
    p[:9] = (np.eye(3) if rotation is None else np.asarray(rotation, np.float64)).reshape(-1)
    p[9:12] = 0.0 if translation is None else np.asarray(translation, np.float64)
    p[12], p[13] = time_lag, close_radius
    return p


def load_multi_sweeps_raw(points: np.ndarray, info: dict, sweeps_num: int = 9, load_dim: int = 5,
                          pad_empty_sweeps: bool = True, remove_close_points: bool = True, test_mode: bool = False,
                          rng=np.random, read=None, close_radius: float = 1.0) -> RawSweeps:
    """`load_multi_sweeps` without the arithmetic: same sweep choice (same draws from ``rng``), but the sweeps stay raw
    and the per-sweep transform / close-point filter / time channel are described by ``params`` for the device."""
    read = read or (lambda path: np.fromfile(path, dtype=np.float32))
    key = np.ascontiguousarray(points, dtype=np.float32)
    arrays, params = [key], [_segment_params()]
    r = close_radius if remove_close_points else -1.0
    sweeps = info["sweeps"]
    if pad_empty_sweeps and not sweeps:
        arrays += [key] * sweeps_num
        params += [_segment_params(close_radius=r)] * sweeps_num
    else:
        for i in _pick_sweeps(len(sweeps), sweeps_num, test_mode, rng):
            sw = sweeps[i]
            arrays.append(np.ascontiguousarray(read(sw["data_path"]), dtype=np.float32).reshape(-1, load_dim))
            params.append(_segment_params(sw["sensor2lidar_rotation"], sw["sensor2lidar_translation"],
                                          info["timestamp"] - sw["timestamp"] / 1e6, r))
    return RawSweeps(arrays, np.stack(params))


def sweep_merge(points: torch.Tensor, seg_offsets: torch.Tensor, seg_params: torch.Tensor):
    """points [N, C] float32 CUDA (raw segments back to back), seg_offsets [S+1] int32 CUDA, seg_params [S, 16] float64
    CUDA -> (merged points [N, C], first ``out_off[-1]`` rows valid; out_off [S+1] int32 CUDA).  No host sync."""
    L.require_cuda(points, "points")
    if points.dtype != torch.float32 or seg_offsets.dtype != torch.int32 or seg_params.dtype != torch.float64:
        raise RuntimeError("sweep_merge: points float32, seg_offsets int32, seg_params float64")
    points, seg_params = points.contiguous(), seg_params.contiguous()
    n, stride = points.shape
    n_seg = seg_offsets.numel() - 1
    if seg_params.shape != (n_seg, 16):
        raise RuntimeError(f"sweep_merge: seg_params {tuple(seg_params.shape)} for {n_seg} segments")
    dev = points.device
    out = torch.empty_like(points)
    out_off = torch.empty(n_seg + 1, dtype=torch.int32, device=dev)
    n_tmp = (n + 1023) // 1024 + 1
    tmp = torch.empty(n_tmp, dtype=torch.int32, device=dev)
    L.run("sweep_merge", L.ptr(points), n, stride, L.ptr(seg_offsets), n_seg, L.ptr(seg_params), L.ptr(out),
          L.ptr(out_off), L.ptr(tmp), C.c_int64(n_tmp), L.stream_ptr(dev))
    return out, out_off


class NuScenesSSLIndex:
    """The `nuscenes_ssl_infos_*.pkl` index as `NuScenesDatasetSSL` reads it (datasets/nuscenes_ssl_dataset.py:176-206,
    228-235): `{'infos': [...], 'metadata': {'version': ...}}`, entries sorted by timestamp, every `load_interval`-th
    kept; an entry yields the loader inputs (`pts_filename`, `sweeps`, `timestamp` in seconds)."""

    def __init__(self, ann_file: str, load_interval: int = 1):
        import pickle
        with open(ann_file, "rb") as f:
            data = pickle.load(f)
        infos = list(sorted(data["infos"], key=lambda e: e["timestamp"]))
        self.data_infos = infos[::load_interval]
        self.metadata = data["metadata"]
        self.version = self.metadata["version"]

    def __len__(self):
        return len(self.data_infos)

    def get_data_info(self, index: int) -> dict:
        info = self.data_infos[index]
        return dict(sample_idx=info["token"], pts_filename=info["lidar_path"], sweeps=info["sweeps"],
                    timestamp=info["timestamp"] / 1e6)

    def load_frame_raw(self, index: int, sweeps_num: int = 9, test_mode: bool = False, rng=np.random) -> RawSweeps:
        """The same sample as ``load_frame`` (same sweep draws), left raw for the device-side merge."""
        info = self.get_data_info(index)
        key = read_points_bin(info["pts_filename"], 5, 5)
        return load_multi_sweeps_raw(key, info, sweeps_num=sweeps_num, test_mode=test_mode, rng=rng)

    def load_frame(self, index: int, sweeps_num: int = 9, test_mode: bool = False, rng=np.random) -> np.ndarray:
        """Raw multi-sweep frame [N, 5] float32 of entry `index` (the first two stages of the train pipeline)."""
        info = self.get_data_info(index)
        key = read_points_bin(info["pts_filename"], 5, 5)
        return load_multi_sweeps(key, info, sweeps_num=sweeps_num, test_mode=test_mode, rng=rng)
