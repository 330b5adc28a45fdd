"""Host-side time per C-ABI entry point of one training step (tiny frames: the GPU is never the limit), i.e. what the
CPU spends inside each geomae_* call (launches, tensor-map encodes, stream events) — complements cpu_time_step.py."""
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200  # noqa: E402,F401
from geomae_b200 import lib as L  # noqa: E402
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402

dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
model = build_model(cfg.model).to(dev).train()
model.set_impl("tc1")
tr = FlatTrainer(model)
frames = [torch.from_numpy(make_frame(s + 1, point_scale=0.02)).to(dev) for s in range(4)]
for _ in range(3):
    tr.train_step(frames)
torch.cuda.synchronize()
acc, cnt = collections.Counter(), collections.Counter()
orig = L.run


def timed(what, *args):
    t = time.perf_counter()
    orig(what, *args)
    acc[what] += time.perf_counter() - t
    cnt[what] += 1


L.run = timed
import geomae_b200.sst, geomae_b200.dense, geomae_b200.voxel, geomae_b200.voxel_encoder, geomae_b200.detector, geomae_b200.windows, geomae_b200.train  # noqa: E402,E401
N = 20
t0 = time.perf_counter()
for _ in range(N):
    tr.train_step(frames)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"enqueue {1e3 * (t1 - t0) / N:.2f} ms/step; inside C-ABI calls {1e3 * sum(acc.values()) / N:.2f} ms/step")
for k, v in acc.most_common(14):
    print(f"  {k:28s} {1e6 * v / N:8.1f} us/step  ({cnt[k] // N} calls)")
