"""Micro-benchmark of the SRA kernels on a realistic 4-frame decoder/encoder token set (CUDA events)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geomae_b200 import lib as L  # noqa: E402
from geomae_b200.dense import layernorm_bwd, tc_linear, tc_wgrad  # noqa: E402
from geomae_b200.sst import _attn_bwd, _attn_fwd  # noqa: E402
from geomae_b200.synthetic import make_frame  # noqa: E402
from geomae_b200.voxel import VoxelGeometry, scatter_frames  # noqa: E402
from geomae_b200.windows import WindowLayout, WindowSpec, pos_table  # noqa: E402

dev = torch.device("cuda:0")
geom = VoxelGeometry((-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (0.256, 0.256, 8), (0.128, 0.128, 2), (0.064, 0.064, 1),
                     (4, 2, 2), (8, 4, 4))
pb = scatter_frames(geom, [torch.from_numpy(make_frame(s + 1)).to(dev) for s in range(4)])
v = pb.n_pillars
spec = WindowSpec((12, 12), [(0, 0), (6, 6)])
perm = torch.randperm(v, device=dev)
sets = {"dec": perm, "enc": perm[: int(v * 0.3)]}
which = sys.argv[1:] or ["attn", "linear", "wgrad", "ln"]


def timeit(fn, n=10):
    """Mean device time (us) of the hand-written kernels launched by fn, from CUPTI kernel records."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    tot = sum(e.device_time_total for e in prof.key_averages() if "k_" in e.key and "geomae" not in e.key)
    return tot / n


for name, rows in sets.items():
    lay = WindowLayout.from_pillars(spec, pb, rows)
    n = rows.shape[0]
    win = lay.shift(0)
    nw = int(lay.n_windows[0])
    ln = torch.diff(lay.win_ptr[0, :nw + 1]).double()
    print(f"[{name}] tokens {n} windows {nw} meanL {ln.mean():.1f} maxL {int(ln.max())} sumL2 {int((ln*ln).sum())}")
    qkv = torch.randn(n, 384, device=dev)
    x = torch.randn(n, 128, device=dev)
    d128 = torch.randn(n, 128, device=dev)
    u = torch.randn(n, 256, device=dev)
    table = pos_table((12, 12), 128, 10000, dev)
    if "attn" in which:
        out, lse = _attn_fwd(qkv, win, 8)
        print(f"  attn fwd {timeit(lambda: _attn_fwd(qkv, win, 8)):8.1f} us   bwd {timeit(lambda: _attn_bwd(qkv, out, lse, d128, win, 8)):8.1f} us")
        print(f"  attn (bf16 mma) fwd {timeit(lambda: _attn_fwd(qkv, win, 8, True)):8.1f} us   bwd {timeit(lambda: _attn_bwd(qkv, out, lse, d128, win, 8, True)):8.1f} us")
    for prec in (1, 3):
        if "linear" in which:
            W384, b384 = torch.randn(384, 128, device=dev) * 0.1, torch.randn(384, device=dev)
            W128, b128 = torch.randn(128, 128, device=dev) * 0.1, torch.randn(128, device=dev)
            W1, W2 = torch.randn(256, 128, device=dev) * 0.1, torch.randn(128, 256, device=dev) * 0.1
            g = torch.ones(128, device=dev)
            t = [timeit(lambda: tc_linear(x, W384, n_out=384, bias=b384, pos_table=table, tok_cell=win["tok_cell"], pos_slabs=2, precision=prec)),
                 timeit(lambda: tc_linear(x, W128, n_out=128, bias=b128, add_src=x, ln=(g, g, 1e-5, True), precision=prec)),
                 timeit(lambda: tc_linear(x, W1, n_out=256, bias=None, precision=prec)),
                 timeit(lambda: tc_linear(u, W2, n_out=128, bias=b128, add_src=x, a_gelu=True, ln=(g, g, 1e-5, True), precision=prec)),
                 timeit(lambda: tc_linear(x, W2, n_out=256, w_mn_major=True, gelu_u=u, precision=prec)),
                 timeit(lambda: tc_linear(qkv, W384, n_out=128, w_mn_major=True, add_src=x, precision=prec))]
            print(f"  p{prec} linear qkv {t[0]:.1f} proj+LN {t[1]:.1f} ffn1 {t[2]:.1f} ffn2+LN {t[3]:.1f} du {t[4]:.1f} dx(K384) {t[5]:.1f} us")
        if "wgrad" in which:
            dW1, db1 = torch.zeros(128, 128, device=dev), torch.zeros(128, device=dev)
            dW2, db2 = torch.zeros(128, 256, device=dev), torch.zeros(128, device=dev)
            dW3, db3 = torch.zeros(384, 128, device=dev), torch.zeros(384, device=dev)
            t = [timeit(lambda: tc_wgrad(d128, x, dW1, db1, precision=prec)),
                 timeit(lambda: tc_wgrad(d128, u, dW2, None, x_gelu=True, precision=prec)),
                 timeit(lambda: tc_wgrad(qkv, x, dW3, db3, pos_table=table, tok_cell=win["tok_cell"], pos_slabs=2, precision=prec))]
            print(f"  p{prec} wgrad 128x128 {t[0]:.1f} 128x256(gelu) {t[1]:.1f} 384x128(pos) {t[2]:.1f} us")
    if "ln" in which:
        st = torch.stack([x.mean(1), torch.rsqrt(x.var(1) + 1e-5)], 1).contiguous()
        dg, db = torch.zeros(128, device=dev), torch.zeros(128, device=dev)
        print(f"  ln_bwd {timeit(lambda: layernorm_bwd(d128, x, st, torch.ones(128, device=dev), dg, db)):.1f} us")
