"""Epoch loop around FlatTrainer — the part of mmcv's EpochBasedRunner + hooks the GeoMAE pre-training run uses
(reference tools/train.py:200-220 -> apis/train.py:35-120; configs/_base_/schedules/cosine_2x.py, default_runtime.py):
cyclic cosine learning rate per iteration, gradient clipping inside the optimiser kernel, `epoch_N.pth` checkpoints
with `state_dict` / `optimizer` / `meta`, resume, and a text log every `log_interval` iterations."""
from __future__ import annotations

import os
import time

import torch

from .train import FlatTrainer, cyclic_lr


def save_checkpoint(path, model, trainer, epoch, it):
    """Same top-level layout as mmcv's save_checkpoint: meta / state_dict / optimizer (the model part loads into the
    reference and into DynamicVoxelNet.load_pretrained by key)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save(dict(meta=dict(epoch=epoch, iter=it, time=time.asctime()),
                    state_dict={k: v.detach().cpu() for k, v in model.state_dict().items()},
                    optimizer=trainer.state_dict()), path)


def load_checkpoint(path, model, trainer=None):
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in ckpt["state_dict"].items()})
    if trainer is not None and "optimizer" in ckpt:
        trainer.load_state_dict(ckpt["optimizer"])
    return ckpt.get("meta", {})


def train(model, loader, work_dir, max_epochs, base_lr=1e-5, trainer: FlatTrainer | None = None, resume_from=None,
          checkpoint_interval=1, log_interval=50, lr_schedule=cyclic_lr, log=print, until_epoch=None):
    """-> (trainer, list of per-iteration losses of this call).  `loader`: a dataset.BatchLoader (or anything with
    set_epoch / __len__ / __iter__ yielding (host_points, augs)).  `until_epoch` stops early (an interrupted run) while
    keeping the learning-rate schedule of the full `max_epochs`."""
    trainer = trainer or FlatTrainer(model, lr=base_lr)
    start_epoch, it = 0, 0
    if resume_from:
        meta = load_checkpoint(resume_from, model, trainer)
        start_epoch, it = meta.get("epoch", 0), meta.get("iter", 0)
        trainer.check_bindings()
    max_iters = max_epochs * len(loader)
    history = []
    model.train()
    for epoch in range(start_epoch, min(max_epochs, until_epoch or max_epochs)):
        loader.set_epoch(epoch)
        t0, pending = time.time(), []
        for i, (host_points, augs) in enumerate(loader):
            loss, _ = trainer.train_step_from_host(host_points, augs=augs, lr=lr_schedule(base_lr, it, max_iters))
            pending.append(loss)
            it += 1
            if (i + 1) % log_interval == 0 or i + 1 == len(loader):
                vals = torch.stack(pending).tolist()            # one read per log line, not per step
                history += vals
                pending = []
                log(f"Epoch [{epoch + 1}][{i + 1}/{len(loader)}] lr: {lr_schedule(base_lr, it - 1, max_iters):.3e}, "
                    f"loss: {sum(vals) / len(vals):.4f}, time: {(time.time() - t0) / (i + 1):.4f} s/iter")
        if (epoch + 1) % checkpoint_interval == 0 or epoch + 1 in (max_epochs, until_epoch):
            save_checkpoint(os.path.join(work_dir, f"epoch_{epoch + 1}.pth"), model, trainer, epoch + 1, it)
    return trainer, history
