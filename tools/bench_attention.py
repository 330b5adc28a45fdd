"""CUDA-event timing of the bf16 attention kernels exactly as the fused path calls them (bf16 q|k|v in, bf16 out;
backward with bf16 dO / dqkv and the precomputed D), on the encoder- and decoder-sized token sets of a 4-frame batch.
GEOMAE_ATTN_TQ=32|64 forces the query-tile size."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geomae_b200 import lib as L
from geomae_b200.synthetic import make_frame
from geomae_b200.voxel import VoxelGeometry, scatter_frames
from geomae_b200.windows import WindowLayout, WindowSpec
dev = torch.device("cuda:0")
preset = sys.argv[1] if len(sys.argv) > 1 else "nuscenes"
geom = VoxelGeometry((-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.256, 0.256, 8), (0.128, 0.128, 2), (0.064, 0.064, 1), (4, 2, 2), (8, 4, 4))
pb = scatter_frames(geom, [torch.from_numpy(make_frame(s + 1, sweeps=int(sys.argv[2]) if len(sys.argv) > 2 else 1)).to(dev) for s in range(4)])
v = pb.n_pillars
spec = WindowSpec((12, 12), [(0, 0), (6, 6)])
perm = torch.randperm(v, device=dev)
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
for name, rows in (("dec", perm), ("enc", perm[: int(v * 0.3)])):
    lay = WindowLayout.from_pillars(spec, pb, rows)
    n = rows.shape[0]
    for shift in (0, 1):
        win = lay.shift(shift)
        qkv = torch.randn(n, 384, device=dev).bfloat16()
        out = torch.empty(n, 128, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(n, 8, device=dev)
        dout = torch.randn(n, 128, device=dev).bfloat16()
        dd = torch.randn(n, 8, device=dev)
        dqkv = torch.empty_like(qkv)
        s = L.stream_ptr(dev)
        def fwd(): L.run("sra_attention_tc_fwd", L.ptr(qkv), n, 8, L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]), L.ptr(win["tok_win"]), L.ptr(out), L.ptr(lse), 1 | 8, s)
        def bwd(): L.run("sra_attention_tc_bwd", L.ptr(qkv), L.ptr(out), L.ptr(lse), L.ptr(dout), n, 8, L.ptr(win["win_ptr"]), L.ptr(win["win_tok"]), L.ptr(win["tok_win"]), L.ptr(dqkv), L.ptr(dd), 1 | 2 | 4, s)
        res = []
        for fn in (fwd, bwd):
            for _ in range(3): fn()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            res.append(sorted(ts)[len(ts) // 2])
        print(f"[{name} shift {shift}] tokens {n}: fwd {res[0]:.1f} us  bwd {res[1]:.1f} us (L2 flushed before each launch)")
