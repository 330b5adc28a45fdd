"""Host-side cost of one training step: wall time of the Python call with the GPU made irrelevant
(tiny frames so kernels are short) vs. full-size frames."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200  # noqa: E402,F401
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402

dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
model = build_model(cfg.model).to(dev).train()
model.set_impl("tc1")
tr = FlatTrainer(model)
for scale in (0.02, 1.0):
    frames = [torch.from_numpy(make_frame(s + 1, point_scale=scale)).to(dev) for s in range(4)]
    for _ in range(3):
        tr.train_step(frames)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        tr.train_step(frames)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"point_scale {scale}: enqueue {1e3*(t1-t0)/10:.2f} ms/step, with final sync {1e3*(t2-t0)/10:.2f} ms/step")
import cProfile, pstats
frames = [torch.from_numpy(make_frame(s + 1, point_scale=0.02)).to(dev) for s in range(4)]
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    tr.train_step(frames)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
