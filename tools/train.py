"""python tools/train.py CONFIG --work-dir DIR [--resume-from CKPT] [--max-epochs N] — single-process entry point in the
shape of the reference's tools/train.py (one process per GPU under torch.distributed.run for multi-GPU)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geomae_b200  # noqa: E402,F401
from geomae_b200.dataset import BatchLoader, build_dataset  # noqa: E402
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.runner import train  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--work-dir", required=True)
    ap.add_argument("--resume-from")
    ap.add_argument("--max-epochs", type=int)
    ap.add_argument("--impl", default="tc1", choices=("tc1", "tc3"))
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    cfg = Config.fromfile(args.config)
    model = build_model(cfg.model).to(f"cuda:{local}")
    model.set_impl(args.impl)
    dataset = build_dataset(cfg.data["train"])
    loader = BatchLoader(dataset, cfg.data["samples_per_gpu"], rank, world, workers=cfg.data.get("workers_per_gpu", 4))
    opt = cfg.optimizer
    trainer = FlatTrainer(model, lr=opt["lr"], betas=tuple(opt.get("betas", (0.9, 0.999))),
                          weight_decay=opt.get("weight_decay", 0.0),
                          max_grad_norm=cfg.optimizer_config["grad_clip"]["max_norm"])
    epochs = args.max_epochs or cfg.runner["max_epochs"]
    train(model, loader, args.work_dir, epochs, base_lr=opt["lr"], trainer=trainer, resume_from=args.resume_from,
          log=(print if rank == 0 else (lambda *_: None)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
