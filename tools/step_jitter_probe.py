"""Per-step host time of a free-running loop (no per-step synchronise): median / p90 / max, and for the slowest steps how
much of it was spent blocked in the input stage (scatter + totals read) vs enqueueing the rest.
python tools/step_jitter_probe.py [steps] [repeats]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import geomae_b200  # noqa: E402,F401
import geomae_b200.detector as D  # noqa: E402
from geomae_b200.registry import Config, build_model  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.train import FlatTrainer  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 40
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
cfg = Config.fromfile(os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py"))
model = build_model(cfg.model).to(dev).train()
model.set_impl("tc1")
batches = [[torch.from_numpy(make_frame(10 * b + s + 1)).to(dev) for s in range(4)] for b in range(8)]
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
acc = [0.0]
orig = D.scatter_frames


def timed_scatter(*a, **k):
    t = time.perf_counter()
    out = orig(*a, **k)
    acc[0] += time.perf_counter() - t
    return out


D.scatter_frames = timed_scatter
import geomae_b200.voxel as V  # noqa: E402
parts = {"run": 0.0, "sizes": 0.0}
_run, _sizes = V.PillarBatch.run, V.PillarBatch.sizes


def t_run(self, *a, **k):
    t = time.perf_counter()
    out = _run(self, *a, **k)
    parts["run"] += time.perf_counter() - t
    return out


def t_sizes(self):
    t = time.perf_counter()
    out = _sizes(self)
    parts["sizes"] += time.perf_counter() - t
    return out


V.PillarBatch.run, V.PillarBatch.sizes = t_run, t_sizes
PRIO = int(os.environ.get("PROBE_PRIORITY", "0"))      # -1: high-priority input stream (shows the stalls)
tr = FlatTrainer(model)
tr.__dict__["_input_stream"] = torch.cuda.Stream(dev, priority=PRIO)
for i in range(24):
    tr.train_step(batches[i % 8])
torch.cuda.synchronize()
import gc
for rep in range(R):
    gc.collect()
    gc.disable()
    host, inp = [], []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        flush.zero_()
        acc[0] = 0.0
        parts["run"] = parts["sizes"] = 0.0
        t = time.perf_counter()
        tr.train_step(batches[i % 8])
        host.append(time.perf_counter() - t)
        inp.append(acc[0])
        if acc[0] > 5e-3:
            print(f"   step {i}: input stage {1e3 * acc[0]:.1f} ms = launch {1e3 * parts['run']:.2f} + totals read "
                  f"{1e3 * parts['sizes']:.2f} + rest (cat, allocations, hand-over) "
                  f"{1e3 * (acc[0] - parts['run'] - parts['sizes']):.2f}")
    e1.record()
    torch.cuda.synchronize()
    gc.enable()
    host, inp = np.array(host) * 1e3, np.array(inp) * 1e3
    worst = np.argsort(-host)[:4]
    print(f"rep {rep}: device {e0.elapsed_time(e1) / K:.3f} ms/step; host median {np.median(host):.2f} p90 "
          f"{np.percentile(host, 90):.2f} max {host.max():.2f} ms; input stage median {np.median(inp):.2f} max {inp.max():.2f}; "
          f"slowest steps (host, input): {[(round(host[j], 2), round(inp[j], 2)) for j in worst]}; "
          f"mem reserved {torch.cuda.memory_reserved() >> 20} MiB, mallocs {torch.cuda.memory_stats()['num_device_alloc']}")
