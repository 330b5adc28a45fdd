"""Drives tools/probe/scatter_probe.so: device time of the point-pass variants on a large batch (CUDA events)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geomae_b200 import lib as L  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.voxel import VoxelGeometry  # noqa: E402

NAMES = {0: "stream only (4x128b/thread)", 1: "tile load only", 2: "+frame search", 3: "+coords/cell",
         4: "+bitmap volatile look-up", 5: "+atomicOr (= k_mark)", 6: "k_mark with ld.cg look-up",
         7: "unconditional atomicOr, no look-up", 8: "register path, volatile", 9: "register path, ld.cg",
         10: "register path, ld.cg, no frame search"}

probe = C.CDLL(os.path.join(ROOT, "tools", "probe", "scatter_probe.so"))
dev = torch.device("cuda:0")
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 256
geom = VoxelGeometry((-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.256, 0.256, 8), (0.128, 0.128, 2), (0.064, 0.064, 1),
                     (4, 2, 2), (8, 4, 4))
base = [torch.from_numpy(make_frame(s + 1)).to(dev) for s in range(8)]
frames = [base[i % 8] for i in range(n_frames)]
pts = torch.cat(frames).contiguous()
offs = [0]
for f in frames:
    offs.append(offs[-1] + f.shape[0])
off = torch.tensor(offs, dtype=torch.int32, device=dev)
n = pts.shape[0]
bitmap = torch.zeros((n_frames * 160000 + 31) // 32, dtype=torch.int32, device=dev)
sink = torch.zeros(4, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = C.c_void_p(L.stream_ptr(dev))
print(f"frames {n_frames} points {n} ({n * 20 / 1e6:.1f} MB)")
for v in sorted(NAMES):
    ts = []
    for it in range(4):
        bitmap.zero_()
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = probe.probe_run(v, C.byref(geom.cstruct), C.c_void_p(pts.data_ptr()), C.c_int64(n), 5,
                             C.c_void_p(off.data_ptr()), n_frames, C.c_void_p(bitmap.data_ptr()),
                             C.c_void_p(sink.data_ptr()), stream)
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0, rc
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = min(ts[1:])
    print(f"  v{v:<2d} {t:8.1f} us  {n * 20 / t / 1e3:7.1f} GB/s  {NAMES[v]}")
