"""Kernel-time table of one training step with torch.profiler (cheap alternative to an ncu launch list)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200  # noqa: E402,F401
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402

impl = sys.argv[1] if len(sys.argv) > 1 else "tc1"
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
model = build_model(cfg.model).to(dev).train()
model.set_impl(impl)
tr = FlatTrainer(model)
frames = [torch.from_numpy(make_frame(s + 1)).to(dev) for s in range(4)]
for _ in range(3):
    tr.train_step(frames)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        tr.train_step(frames)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by=(sys.argv[2] if len(sys.argv) > 2 else "cuda_time_total"), row_limit=45, max_name_column_width=70))
