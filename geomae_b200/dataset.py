"""NuScenesDatasetSSL + batch loader: from `nuscenes_ssl_infos_*.pkl` and `.pcd.bin` sweeps to
the pinned host frames `FlatTrainer.train_step_from_host` consumes.

`NuScenesDatasetSSL` takes the reference's constructor arguments and pipeline list unchanged
(datasets/nuscenes_ssl_dataset.py:16-140, configs/mae_sst/…6x_1e-5.py:167-197,257-268) and splits the pipeline where
the work moves to the device: the file stages (LoadPointsFromFile, LoadPointsFromMultiSweeps) run on the host exactly
as the reference's numpy code does (data.py), the geometric stages (GlobalRotScaleTrans, RandomFlip3D,
PointsRangeFilter) only DRAW their parameters here, in the reference's order, and are applied by
`geomae_augment_filter` in front of the scatter; PointShuffle / DefaultFormatBundle3D / Collect3D have nothing left
to do (everything downstream of the scatter is order-free, the frames stay plain float32 arrays).

`BatchLoader` restates the batching of mmdet 2.20's `DistributedGroupSampler` + `DataLoader(collate)` as the
reference's `build_dataloader` configures them (third-party code, absent from the reference tree: epoch-seeded
permutation, padded to a multiple of samples_per_gpu x world size, batches permuted, one contiguous share per rank) and
reads ahead on a few host threads (numpy file reads release the GIL)."""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .data import (Augmentation, NuScenesSSLIndex, RawSweeps, draw_augmentation, load_multi_sweeps, load_multi_sweeps_raw,
                   read_points_bin)
from .registry import Registry

DATASETS = Registry("dataset")
_NO_OPS = ("PointShuffle", "DefaultFormatBundle3D", "Collect3D")


@DATASETS.register_module()
class NuScenesDatasetSSL:
    CLASSES = ("car", "truck", "trailer", "bus", "construction_vehicle", "bicycle", "motorcycle", "pedestrian",
               "traffic_cone", "barrier")

    def __init__(self, ann_file, pipeline=None, data_root=None, classes=None, load_interval=1, with_velocity=True,
                 modality=None, box_type_3d="LiDAR", filter_empty_gt=False, test_mode=False,
                 eval_version="detection_cvpr_2019", use_valid_flag=False, device_merge=False):
        """``device_merge`` (not a reference argument): leave the sweeps raw (``data.RawSweeps``) so that the sweep ->
        key-frame transform, the close-point filter and the concatenation run on the device (geomae_sweep_merge)
        instead of in numpy on the loader threads; same sweep choice, same result."""
        self.data_root, self.ann_file, self.test_mode, self.device_merge = data_root, ann_file, test_mode, device_merge
        self.CLASSES = tuple(classes) if classes is not None else self.CLASSES
        self.modality = modality or dict(use_camera=False, use_lidar=True, use_radar=False, use_map=False,
                                         use_external=False)
        self.index = NuScenesSSLIndex(ann_file, load_interval)
        self.data_infos, self.version = self.index.data_infos, self.index.version
        self.flag = np.zeros(len(self), dtype=np.uint8)        # one aspect-ratio group: there are no images (custom_3d.py:357-364)
        self.load = dict(load_dim=5, use_dim=5)
        self.sweeps = None
        self.aug = dict(rot_range=None, scale_ratio_range=None, flip_h=None, flip_v=None)
        self.point_cloud_range = None
        for step in pipeline or []:
            kind = step["type"]
            if kind == "LoadPointsFromFile":
                assert step.get("coord_type", "LIDAR") == "LIDAR"
                self.load = dict(load_dim=step.get("load_dim", 6), use_dim=step.get("use_dim", [0, 1, 2]))
            elif kind == "LoadPointsFromMultiSweeps":
                self.sweeps = dict(sweeps_num=step.get("sweeps_num", 10), load_dim=step.get("load_dim", 5),
                                   use_dim=tuple(step.get("use_dim", [0, 1, 2, 4])),
                                   pad_empty_sweeps=step.get("pad_empty_sweeps", False),
                                   remove_close_points=step.get("remove_close", False),
                                   test_mode=step.get("test_mode", False))
            elif kind == "GlobalRotScaleTrans":
                if any(float(s) != 0.0 for s in step.get("translation_std", [0, 0, 0])):
                    raise NotImplementedError("translation noise is 0 in every GeoMAE config")
                self.aug.update(rot_range=tuple(step["rot_range"]), scale_ratio_range=tuple(step["scale_ratio_range"]))
            elif kind == "RandomFlip3D":
                self.aug.update(flip_h=step.get("flip_ratio_bev_horizontal", 0.0),
                                flip_v=step.get("flip_ratio_bev_vertical", 0.0))
            elif kind == "PointsRangeFilter":
                self.point_cloud_range = tuple(step["point_cloud_range"])
            elif kind not in _NO_OPS:
                raise NotImplementedError(f"pipeline step {kind} is not part of the GeoMAE pre-training pipeline")

    def __len__(self):
        return len(self.index)

    def get_data_info(self, index):
        info = self.index.get_data_info(index)
        if self.data_root and not os.path.isabs(info["pts_filename"]) and not os.path.exists(info["pts_filename"]):
            info["pts_filename"] = os.path.join(self.data_root, info["pts_filename"])
        return info

    def draw(self, rng=np.random) -> Augmentation:
        """The random stages' draws for one sample, in pipeline order (identity where a stage is absent)."""
        a = self.aug
        if a["rot_range"] is None and a["flip_h"] is None:
            return Augmentation()
        return draw_augmentation(rng, a["rot_range"] or (0.0, 0.0), a["scale_ratio_range"] or (1.0, 1.0),
                                 a["flip_h"] or 0.0, a["flip_v"] or 0.0)

    def __getitem__(self, index, rng=np.random):
        """-> dict(points [N, C] float32 raw multi-sweep frame, aug = this sample's draws, sample_idx)."""
        info = self.get_data_info(index)
        pts = read_points_bin(info["pts_filename"], self.load["load_dim"], self.load["use_dim"])
        if self.sweeps is not None and self.device_merge and tuple(self.sweeps["use_dim"]) == tuple(range(pts.shape[1])):
            kw = {k: v for k, v in self.sweeps.items() if k != "use_dim"}
            return dict(points=load_multi_sweeps_raw(pts, info, rng=rng, **kw), aug=self.draw(rng),
                        sample_idx=info["sample_idx"])
        if self.sweeps is not None:
            pts = load_multi_sweeps(pts, info, rng=rng, **self.sweeps)
        return dict(points=np.ascontiguousarray(pts, dtype=np.float32), aug=self.draw(rng), sample_idx=info["sample_idx"])


def build_dataset(cfg, default_args=None):
    return DATASETS.build(cfg, default_args)


def epoch_indices(n, samples_per_gpu, rank, world, epoch, seed=0, shuffle=True):
    """One rank's sample indices for one epoch (mmdet 2.20 DistributedGroupSampler.__iter__ with a single group):
    permutation seeded by epoch + seed, padded by wrapping around to a multiple of samples_per_gpu * world, the
    batches permuted once more, then the rank's contiguous share."""
    g = torch.Generator()
    g.manual_seed(epoch + seed)
    order = torch.randperm(n, generator=g).tolist() if shuffle else list(range(n))
    chunk = samples_per_gpu * world
    total = int(math.ceil(n / chunk)) * chunk
    extra = total - n
    order = order + order * (extra // n) + order[:extra % n]
    if shuffle:
        order = [order[j] for i in torch.randperm(total // samples_per_gpu, generator=g).tolist()
                 for j in range(i * samples_per_gpu, (i + 1) * samples_per_gpu)]
    share = total // world
    return order[rank * share:(rank + 1) * share]


class BatchLoader:
    """Iterates one epoch of (host_points, augs) batches of a rank: ``host_points`` a list of samples_per_gpu pinned
    float32 tensors, ``augs`` their augmentation draws.  ``set_epoch`` as torch's samplers have it."""

    def __init__(self, dataset, samples_per_gpu, rank=0, world=1, shuffle=True, seed=0, workers=4, prefetch=2,
                 pin_memory=True):
        self.dataset, self.spg, self.rank, self.world = dataset, samples_per_gpu, rank, world
        self.shuffle, self.seed, self.workers, self.prefetch, self.pin = shuffle, seed, workers, prefetch, pin_memory
        self.epoch = 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        return len(epoch_indices(len(self.dataset), self.spg, self.rank, self.world, 0, shuffle=False)) // self.spg

    def _sample(self, index, seed):
        rng = np.random.RandomState(seed)        # per-sample stream: the result does not depend on thread timing
        item = self.dataset.__getitem__(index, rng=rng)
        pin = (lambda a: torch.from_numpy(a).pin_memory()) if self.pin and torch.cuda.is_available() else torch.from_numpy
        pts = item["points"]
        if isinstance(pts, RawSweeps):
            seen = {}          # padded samples repeat the key frame: pin it once
            return RawSweeps([seen.setdefault(id(a), pin(a)) for a in pts.arrays], pts.params), item["aug"]
        return pin(pts), item["aug"]

    def __iter__(self):
        idx = epoch_indices(len(self.dataset), self.spg, self.rank, self.world, self.epoch, self.seed, self.shuffle)
        base = (self.seed * 1000003 + self.epoch) * 1000003
        batches = [idx[i:i + self.spg] for i in range(0, len(idx), self.spg)]
        with ThreadPoolExecutor(max_workers=max(1, self.workers)) as pool:
            def submit(b):
                return [pool.submit(self._sample, j, (base + self.rank * len(idx) + b * self.spg + k) % (2 ** 32))
                        for k, j in enumerate(batches[b])]
            ahead = [submit(b) for b in range(min(self.prefetch, len(batches)))]
            for b in range(len(batches)):
                futs = ahead.pop(0)
                if b + self.prefetch < len(batches):
                    ahead.append(submit(b + self.prefetch))
                got = [f.result() for f in futs]
                yield [g[0] for g in got], [g[1] for g in got]
