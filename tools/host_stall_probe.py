"""Where the host waits inside a free-running training loop: time spent in scatter_frames (the input stage incl. the
blocking read of the pillar totals) vs. the rest of the step's enqueue, for the input stage on the compute stream, on
its own stream, and on its own high-priority stream.  python tools/host_stall_probe.py [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200  # noqa: E402,F401
import geomae_b200.detector as D  # noqa: E402
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
model = build_model(cfg.model).to(dev).train()
model.set_impl("tc1")
batches = [[torch.from_numpy(make_frame(10 * b + s + 1)).to(dev) for s in range(4)] for b in range(4)]
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
acc = [0.0]
orig = D.scatter_frames


def timed_scatter(*a, **k):
    t = time.perf_counter()
    out = orig(*a, **k)
    acc[0] += time.perf_counter() - t
    return out


D.scatter_frames = timed_scatter
for mode in ("compute stream", "own stream", "own stream, high priority"):
    tr = FlatTrainer(model, overlap_input=mode != "compute stream")
    if mode.endswith("priority"):
        tr.__dict__["_input_stream"] = torch.cuda.Stream(dev, priority=-1)
    for i in range(6):
        tr.train_step(batches[i % 4])
    torch.cuda.synchronize()
    acc[0] = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        flush.zero_()
        tr.train_step(batches[i % 4])
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{mode:28s}: device {e0.elapsed_time(e1) / K:.3f} ms/step, host enqueue {1e3 * host / K:.3f} ms/step, "
          f"of which input stage {1e3 * acc[0] / K:.3f} ms")
