"""Soak: N host-fed training steps with fresh random augmentations; reports throughput per 200 steps, reserved memory
and cudaMalloc count (both must level off) and that the loss stays finite / goes down."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200 as G
from geomae_b200.data import draw_augmentation
from geomae_b200.registry import Config
from geomae_b200.synthetic import make_frame
from geomae_b200.train import FlatTrainer, cyclic_lr
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
torch.manual_seed(0)
model = G.build_detector(cfg.model).to(dev).train(); model.set_impl("tc1")
tr = FlatTrainer(model, lr=1e-5)
host = [[torch.from_numpy(make_frame(100 * b + s)).pin_memory() for s in range(4)] for b in range(16)]
rs = np.random.RandomState(0)
pending, t0 = [], time.perf_counter()
for i in range(N):
    batch = host[i % 16]
    loss, _ = tr.train_step_from_host(batch, augs=[draw_augmentation(rs) for _ in batch], lr=cyclic_lr(1e-5, i, N))
    pending.append(loss)
    if (i + 1) % 200 == 0:
        vals = torch.stack(pending).tolist(); pending = []
        dt = time.perf_counter() - t0; t0 = time.perf_counter()
        st = torch.cuda.memory_stats()
        print(f"steps {i - 198:4d}-{i + 1:4d}: {200 * 4 / dt:7.1f} frames/s, loss {np.mean(vals):.4f} (finite {bool(np.isfinite(vals).all())}), "
              f"reserved {torch.cuda.memory_reserved() >> 20} MiB, cudaMallocs {st['num_device_alloc']}", flush=True)
