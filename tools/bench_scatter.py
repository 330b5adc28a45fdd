"""Per-kernel device time of the voxel-scatter and geometric-target stages on a large batch (CUPTI records)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.voxel import VoxelGeometry, scatter_frames  # noqa: E402

dev = torch.device("cuda:0")
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
geom = VoxelGeometry((-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.256, 0.256, 8), (0.128, 0.128, 2), (0.064, 0.064, 1),
                     (4, 2, 2), (8, 4, 4))
base = [torch.from_numpy(make_frame(s + 1)).to(dev) for s in range(8)]
frames = [base[i % 8] for i in range(n_frames)]
pb = scatter_frames(geom, frames)
v, vm, vl = pb.sizes()
p = pb.points.shape[0]
print(f"frames {n_frames} points {p} pillars {v} med {vm} low {vl}")
for _ in range(2):
    pb.run()
    pb.geom_targets()
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        pb.run()
        pb.geom_targets()
    torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"  {e.device_time_total / 5:9.1f} us  x{e.count // 5:2d}  {e.key[:90]}")
        tot += e.device_time_total / 5
b_sc = 24.0 * p + 32.0 * v + 20.0 * (vm + vl)
b_gt = 20.0 * vm + 28.0 * v + 24.0 * v
print(f"total {tot:.1f} us; algorithmic bytes scatter {b_sc/1e6:.1f} MB, targets {b_gt/1e6:.1f} MB")
